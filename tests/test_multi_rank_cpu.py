"""-m "not gpu": the N > 1 scheme on CPU, world_size 2 over gloo. Each rank traces its share of every global frame's
samples (rank r: samples r*spp .. (r+1)*spp-1, the partition rlpt_config.rank/world_size selects), the Q accumulators
are all-reduced through the same hook contract the library uses (rlpt/dist.py), and every rank applies the same merge.
Compute here is the oracle (the library itself needs a GPU; its 2-GPU twin is tests/test_gpu_parity.py::test_two_gpus*).
Checks: the union over ranks equals one rank tracing all samples -- accumulators, merged Q, CDFs, image."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

W = H = 24
SPP = 2


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "reinforcement-light-rays-pathtracer_b200"))
    import torch.distributed as dist
    from checkers import Oracle
    from rlpt.dist import host_allreduce_hook
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "scenes.npz"))
    s = {k.split("/")[1]: z[k] for k in z.files if k.startswith("cornell/")}
    orc = Oracle(); orc.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"]); orc.rmap_build(0.01); orc.rmap_update_distributions(); orc.rmap_merge_frame()
    hook = host_allreduce_hook()
    img = np.zeros((W * H, 3), np.float32)
    for f in range(2):
        o, _ = orc.render_frame(1, W, H, SPP, sample0=(f * world + rank) * SPP, fma_mode=1, td_mode=1)
        img += o
        acc, cnt = orc.rmap_acc()
        acc32, cnt32 = acc.astype(np.float32), cnt.astype(np.int32)               # the library's buffers: float32 sums, uint32 counts
        hook(acc32.ctypes.data, acc32.size, 0, 0); hook(cnt32.ctypes.data, cnt32.size, 1, 0)
        orc.rmap_set_acc(acc32.astype(np.float64), cnt32.astype(np.uint32))
        orc.rmap_merge_frame(); orc.rmap_update_distributions()
    hook(img.ctypes.data, img.size, 0, 0)                                          # rlpt_frame_allreduce
    q, cdf, vis, irr = orc.rmap_state()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), q=q, cdf=cdf, vis=vis, img=img)
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path, oracle, golden_scenes):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("q", "cdf", "vis", "img"):
        assert np.array_equal(r0[k], r1[k]), k                                    # replicas stay bit-identical
    # single rank tracing the same global frames (2*SPP samples each)
    s = golden_scenes["cornell"]
    oracle.scene_set(s["sv"], s["srgb"], s["lv"], s["lrgb"]); oracle.rmap_build(0.01); oracle.rmap_update_distributions(); oracle.rmap_merge_frame()
    img = np.zeros((W * H, 3), np.float32)
    for f in range(2):
        o, _ = oracle.render_frame(1, W, H, 2 * SPP, sample0=f * 2 * SPP, fma_mode=1, td_mode=1)
        img += o
        oracle.rmap_merge_frame(); oracle.rmap_update_distributions()
    q, cdf, vis, irr = oracle.rmap_state()
    assert np.array_equal(vis, r0["vis"]) and int(vis.sum()) > 0
    assert np.allclose(q, r0["q"], rtol=1e-5, atol=1e-7)                           # float32 partial sums vs one double sum
    assert np.allclose(cdf, r0["cdf"], rtol=1e-4, atol=1e-6)
    assert np.allclose(img, r0["img"], rtol=1e-4, atol=1e-5)


def test_sample_partition_is_disjoint_and_complete():
    """sample_base = (frame*world + rank)*spp (csrc/rlpt_capi.cu enqueue_trace): ranks tile each global frame exactly"""
    for world in (1, 2, 4, 8):
        for spp in (1, 4, 32):
            seen = []
            for frame in range(3):
                for rank in range(world):
                    base = (frame * world + rank) * spp
                    seen += list(range(base, base + spp))
            assert sorted(seen) == list(range(3 * world * spp))
