// scratch/ubench/copy_bw.cu -- per-SM global(L2) -> shared copy bandwidth on B200 for the mechanisms k_dqn_forward could stream its weights with:
//   mode 0: cp.async.bulk (UBLKCP), one lane, `depth` copies of `chunk` bytes in flight
//   mode 1: cp.async 16 B (LDGSTS) from `nthreads` threads, commit/wait groups, `depth` chunks in flight
//   mode 2: plain ld.global.v4 -> st.shared from `nthreads` threads
// All CTAs (one per SM, grid = #SMs or fewer) read the same 320 KB buffer (L2 resident), as the forward kernel's CTAs do.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_bw copy_bw.cu ; run: ./copy_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int STAGES = 8;
__global__ void k_copy(const uint8_t* __restrict__ src, int total_bytes, int chunk, int depth, int iters, int mode, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[STAGES];
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) { for (int i = 0; i < STAGES; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int n_chunks = total_bytes / chunk;
    long long t0 = clock64();
    if (mode == 0) {
        if (t == 0) {
            int issued = 0, done = 0; const int total = iters * n_chunks;
            for (; issued < depth && issued < total; ++issued) { int st = issued % depth; mbar_expect_tx(&bars[st], chunk); bulk_copy(smem + (size_t)st * chunk, src + (size_t)(issued % n_chunks) * chunk, chunk, &bars[st]); }
            for (; done < total; ++done) {
                int st = done % depth; mbar_wait(&bars[st], (done / depth) & 1);
                if (issued < total) { mbar_expect_tx(&bars[st], chunk); bulk_copy(smem + (size_t)st * chunk, src + (size_t)(issued % n_chunks) * chunk, chunk, &bars[st]); ++issued; }
            }
        }
    } else if (mode == 1) {
        const int total = iters * n_chunks; const int per = chunk / 16;
        int issued = 0;
        for (int done = 0; done < total; ++done) {
            for (; issued < done + depth && issued < total; ++issued) {
                const uint8_t* s = src + (size_t)(issued % n_chunks) * chunk; uint8_t* d = smem + (size_t)(issued % depth) * chunk;
                for (int i = t; i < per; i += nt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + 16 * i)), "l"(s + 16 * i) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // wait until the oldest group is complete: at most (issued - done - 1) groups may stay pending
            const int pend = issued - done - 1;
            if (pend >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory"); else if (pend == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else if (pend == 1) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
    } else {
        const int total = iters * n_chunks; const int per = chunk / 16; float acc = 0.f;
        for (int c = 0; c < total; ++c) {
            const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)(c % n_chunks) * chunk); uint4* d = reinterpret_cast<uint4*>(smem + (size_t)(c % depth) * chunk);
            for (int i = t; i < per; i += nt) { uint4 v = __ldg(s + i); d[i] = v; acc += __uint_as_float(v.x); }
        }
        if (acc == 1.2345f) *sink = acc;
    }
    __syncthreads();
    long long t1 = clock64();
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
    int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    const int total_bytes = 320 * 1024;   // (a multiple of every chunk size used below is not required: n_chunks = total / chunk)
    uint8_t* src; cudaMalloc(&src, total_bytes); cudaMemset(src, 1, total_bytes);
    long long* cyc; cudaMalloc(&cyc, sizeof(long long) * 256); float* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int mode, chunk, depth, threads, grid; } cfgs[] = {
        {0, 40960, 1, 32, 0}, {0, 40960, 2, 32, 0}, {0, 40960, 4, 32, 0}, {0, 65536, 1, 32, 0}, {0, 65536, 2, 32, 0}, {0, 65536, 3, 32, 0}, {0, 32768, 3, 32, 0}, {0, 26624, 3, 32, 0}, {0, 16384, 3, 32, 0}, {0, 8192, 3, 32, 0}, {0, 2048, 8, 32, 0},
        {0, 20480, 1, 32, 0}, {0, 20480, 2, 32, 0}, {0, 20480, 3, 32, 0}, {0, 20480, 6, 32, 0}, {0, 10240, 8, 32, 0}, {0, 5120, 8, 32, 0}, {0, 20480, 3, 32, 32}, {0, 20480, 6, 32, 32}, {0, 20480, 6, 32, 1},
        {1, 20480, 2, 128, 0}, {1, 20480, 3, 128, 0}, {1, 20480, 4, 128, 0}, {1, 20480, 4, 256, 0}, {1, 20480, 4, 512, 0}, {1, 20480, 4, 128, 32}, {1, 20480, 4, 256, 1},
        {2, 20480, 2, 256, 0}, {2, 20480, 2, 512, 0}, {2, 20480, 2, 1024, 0},
    };
    for (auto& c : cfgs) {
        const int grid = c.grid ? c.grid : n_sm, iters = 40;
        for (int rep = 0; rep < 2; ++rep) k_copy<<<grid, c.threads, (size_t)c.chunk * c.depth, 0>>>(src, total_bytes, c.chunk, c.depth, iters, c.mode, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
        const double bytes = (double)iters * (total_bytes / c.chunk) * c.chunk;
        printf("mode %d chunk %5d depth %d threads %4d grid %3d: %.1f B/cycle per SM (%s)\n", c.mode, c.chunk, c.depth, c.threads, grid, bytes / avg, cudaGetErrorString(e));
    }
    return 0;
}
