// nccl_hook.cpp -- the all-reduce hook of include/rlpt.h (rlpt_allreduce_fn) on NCCL, for C++ hosts.
// One process per GPU (or one thread per GPU with ncclCommInitAll): create the communicator the usual way, then
//   rlpt_set_allreduce(ctx, rlpt_nccl_allreduce, (void*)comm);
// The collective is enqueued on the library's own stream, so it is ordered between the tracing kernels and the merge
// kernel with no host synchronisation; over NVLink 5 / NVSwitch NCCL picks its NVLS / ring algorithm itself.
#include <nccl.h>
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" int rlpt_nccl_allreduce(void* d_buf, uint64_t count, int dtype, void* cuda_stream, void* comm) {
    ncclDataType_t t = dtype == 0 ? ncclFloat32 : ncclUint32;
    ncclResult_t r = ncclAllReduce(d_buf, d_buf, (size_t)count, t, ncclSum, (ncclComm_t)comm, (cudaStream_t)cuda_stream);
    return r == ncclSuccess ? 0 : (int)r;
}
