"""rlpt -- Python (ctypes) binding of the C ABI in include/rlpt.h.

Used by tests/ and bench.py only; the product is the shared library (csrc/) and the C++ host mirror (host/).
There is no fallback: importing works anywhere (so symbol tests can run on a CPU box), but every call that computes
goes straight to the CUDA library and raises RlptError when that fails (e.g. no GPU).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", os.environ.get("RLPT_LIB_NAME", "librlpt.so"))
CELLS = 144

HIT_NOTHING, HIT_AREA_LIGHT, HIT_SURFACE = 0, 1, 2
TRAVERSAL_AUTO, TRAVERSAL_BVH, TRAVERSAL_BRUTE = 0, 1, 2


class RlptError(RuntimeError):
    pass


class Config(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("spp", ctypes.c_int32), ("max_bounces", ctypes.c_int32),
                ("env_light", ctypes.c_float), ("area_per_sample", ctypes.c_float), ("max_dist", ctypes.c_float),
                ("initial_radiance", ctypes.c_float), ("radiance_threshold", ctypes.c_float), ("seed", ctypes.c_uint32),
                ("traversal", ctypes.c_int32), ("rank", ctypes.c_int32), ("world_size", ctypes.c_int32)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("paths", "path_length_sum", "zero_contribution_paths", "ray_casts", "device_seconds", "frames", "kernel_launches",
                                           "triangle_tests", "box_tests", "trace_seconds", "merge_seconds", "kd_fallbacks",
                                           "isect_seconds", "isect_launches", "shade_seconds", "shade_launches", "tail_seconds", "tail_launches",
                                           "dqn_forward_seconds", "dqn_forward_launches", "dqn_forward_rays", "train_steps", "train_seconds")]


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)

# every symbol include/rlpt.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = [
    "rlpt_last_error", "rlpt_version", "rlpt_ctx_create", "rlpt_ctx_destroy", "rlpt_sync", "rlpt_stream", "rlpt_config_default", "rlpt_config_set",
    "rlpt_config_get", "rlpt_set_allreduce", "rlpt_scene_upload", "rlpt_scene_info", "rlpt_scene_bvh_download", "rlpt_scene_bvh4_info", "rlpt_scene_bvh4_download", "rlpt_neuralq_set_hyper", "rlpt_set_max_direction", "rlpt_radiance_map_build_seconds", "rlpt_camera_set", "rlpt_closest_hit",
    "rlpt_closest_hit_device", "rlpt_radiance_map_build", "rlpt_radiance_map_info", "rlpt_radiance_map_tree", "rlpt_radiance_map_find_closest",
    "rlpt_radiance_map_set_q", "rlpt_radiance_map_update_distributions", "rlpt_radiance_map_download", "rlpt_radiance_map_delta_download",
    "rlpt_radiance_map_save_q", "rlpt_radiance_map_load_q", "rlpt_render_default", "rlpt_render_sarsa", "rlpt_sarsa_trace", "rlpt_sarsa_merge",
    "rlpt_render_sarsa_frozen", "rlpt_frame_reset", "rlpt_frame_allreduce", "rlpt_frame_download", "rlpt_frame_download_argb", "rlpt_frame_save_bmp",
    "rlpt_stats", "rlpt_stats_reset", "rlpt_measure_fp32_peak", "rlpt_capture_rays",
    "rlpt_dqn_set_vertices", "rlpt_dqn_init", "rlpt_dqn_load_text", "rlpt_dqn_save_text", "rlpt_dqn_param_count", "rlpt_dqn_set_params", "rlpt_dqn_get_params",
    "rlpt_dqn_forward", "rlpt_render_pretrained", "rlpt_dqn_train_batch", "rlpt_dqn_train_supervised", "rlpt_render_voronoi", "rlpt_p2p_blob_bytes", "rlpt_p2p_export", "rlpt_p2p_import", "rlpt_p2p_close", "rlpt_dqn_get_grads", "rlpt_render_neuralq", "rlpt_neuralq_last_loss",
]

_lib = None


def lib():
    """Load librlpt.so (built in-tree by build.sh / __graft_entry__.build()). Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RlptError("CUDA library not built: %s is missing (run __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.rlpt_last_error.restype = ctypes.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def default_config(**kw):
    c = Config()
    lib().rlpt_config_default(ctypes.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class Context:
    """One GPU, one stream. Mirrors the order of calls in the reference's main.cu."""

    def __init__(self, device=0, **config):
        self.L = lib()
        self.h = ctypes.c_void_p()
        self._ck(self.L.rlpt_ctx_create(int(device), ctypes.byref(self.h)))
        self.cfg = default_config(**config)
        self._ck(self.L.rlpt_config_set(self.h, ctypes.byref(self.cfg)))
        self._hook = None
        self.n_vol = 0

    def _ck(self, rc):
        if rc != 0:
            raise RlptError("rlpt error %d: %s" % (rc, (self.L.rlpt_last_error() or b"").decode()))

    def close(self):
        if self.h:
            self.L.rlpt_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, **kw):
        for k, v in kw.items():
            setattr(self.cfg, k, v)
        self._ck(self.L.rlpt_config_set(self.h, ctypes.byref(self.cfg)))

    def sync(self):
        self._ck(self.L.rlpt_sync(self.h))

    def stream(self):
        s = ctypes.c_void_p()
        self._ck(self.L.rlpt_stream(self.h, ctypes.byref(s)))
        return s.value

    def set_allreduce(self, fn):
        """fn(ptr:int, count:int, dtype:int(0=f32,1=u32), stream:int) -> None; called on the library's stream order."""
        if fn is None:
            self._hook = None
            self._ck(self.L.rlpt_set_allreduce(self.h, None, None))
            return

        def tramp(ptr, count, dtype, stream, user):
            try:
                fn(ptr, count, dtype, stream)
                return 0
            except Exception as e:  # noqa: BLE001 -- report through the C status
                print("rlpt all-reduce hook failed:", e)
                return 1
        self._hook = ALLREDUCE_FN(tramp)
        self._ck(self.L.rlpt_set_allreduce(self.h, self._hook, None))

    def p2p_export(self):
        blob = ctypes.create_string_buffer(self.L.rlpt_p2p_blob_bytes())
        self._ck(self.L.rlpt_p2p_export(self.h, blob))
        return blob.raw

    def p2p_import(self, blobs):
        """blobs: the p2p_export() results of all ranks in rank order"""
        data = b"".join(blobs)
        self._ck(self.L.rlpt_p2p_import(self.h, data, len(blobs)))

    def p2p_close(self):
        self._ck(self.L.rlpt_p2p_close(self.h))

    # ---- scene / camera
    def scene_upload(self, sv, srgb, lv, lrgb):
        sv, srgb = _f32(sv).reshape(-1, 9), _f32(srgb).reshape(-1, 3)
        lv, lrgb = _f32(lv).reshape(-1, 9), _f32(lrgb).reshape(-1, 3)
        self._ck(self.L.rlpt_scene_upload(self.h, _p(sv), _p(srgb), len(sv), _p(lv), _p(lrgb), len(lv)))

    def scene_info(self):
        v = [ctypes.c_int() for _ in range(4)]
        self._ck(self.L.rlpt_scene_info(self.h, *[ctypes.byref(x) for x in v]))
        return dict(n_surfaces=v[0].value, n_lights=v[1].value, bvh_nodes=v[2].value, bvh_depth=v[3].value)

    def bvh_download(self):
        n = self.scene_info()["bvh_nodes"]
        a = np.zeros((n, 16), np.float32)
        self._ck(self.L.rlpt_scene_bvh_download(self.h, _p(a), n))
        return a

    def bvh4_download(self):
        """the 4-wide tree the kernels walk: (nodes [n, 28] float32, record_gid [n_primitives] int32, depth, leaf_max)"""
        v = [ctypes.c_int() for _ in range(3)]
        self._ck(self.L.rlpt_scene_bvh4_info(self.h, *[ctypes.byref(x) for x in v]))
        info = self.scene_info(); m = info["n_surfaces"] + info["n_lights"]
        a = np.zeros((v[0].value, 28), np.float32); g = np.zeros(m, np.int32)
        self._ck(self.L.rlpt_scene_bvh4_download(self.h, _p(a), v[0].value, _p(g), m))
        return a, g, v[1].value, v[2].value

    def camera_set(self, pos, yaw_y=0.0, yaw_x=0.0):
        p = _f32(list(pos)[:3] + [1.0])
        self._ck(self.L.rlpt_camera_set(self.h, _p(p), ctypes.c_float(yaw_y), ctypes.c_float(yaw_x)))

    # ---- closest hit
    def closest_hit(self, org, dir, traversal=0, count=False):
        org, dir = _f32(org).reshape(-1, 3), _f32(dir).reshape(-1, 3)
        n = len(org)
        ty, ix, t = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        cnt = np.zeros(2, np.uint64) if count else None
        self._ck(self.L.rlpt_closest_hit(self.h, _p(org), _p(dir), n, int(traversal), _p(ty), _p(ix), _p(t), _p(cnt)))
        return (ty, ix, t, cnt) if count else (ty, ix, t)

    def closest_hit_device(self, d_org, d_dir, n, d_type, d_index, d_t, traversal=0, d_counters=None):
        self._ck(self.L.rlpt_closest_hit_device(self.h, ctypes.c_void_p(d_org), ctypes.c_void_p(d_dir), int(n), int(traversal),
                                                ctypes.c_void_p(d_type), ctypes.c_void_p(d_index), ctypes.c_void_p(d_t),
                                                ctypes.c_void_p(d_counters) if d_counters else None))

    # ---- radiance map
    def radiance_map_build(self):
        self._ck(self.L.rlpt_radiance_map_build(self.h))
        a, b = ctypes.c_int(), ctypes.c_int()
        self._ck(self.L.rlpt_radiance_map_info(self.h, ctypes.byref(a), ctypes.byref(b)))
        self.n_vol, self.n_tree = a.value, b.value
        return self.n_vol

    def radiance_map_tree(self):
        nt = self.n_tree
        d = dict(dim=np.zeros(nt, np.int32), leaf=np.zeros(nt, np.int32), left=np.zeros(nt, np.uint32), right=np.zeros(nt, np.uint32),
                 data=np.zeros(nt, np.float32), pos=np.zeros((nt, 3), np.float32), nrm=np.zeros((nt, 3), np.float32))
        self._ck(self.L.rlpt_radiance_map_tree(self.h, *[_p(d[k]) for k in ("dim", "leaf", "left", "right", "data", "pos", "nrm")]))
        return d

    def find_closest(self, pos, nrm):
        pos, nrm = _f32(pos).reshape(-1, 3), _f32(nrm).reshape(-1, 3)
        out = np.zeros(len(pos), np.int32)
        self._ck(self.L.rlpt_radiance_map_find_closest(self.h, _p(pos), _p(nrm), len(pos), _p(out)))
        return out

    def radiance_map_set_q(self, q, visits=None):
        q = _f32(q)
        v = np.ascontiguousarray(visits, dtype=np.uint32) if visits is not None else None
        self._ck(self.L.rlpt_radiance_map_set_q(self.h, _p(q), _p(v)))

    def radiance_map_update_distributions(self):
        self._ck(self.L.rlpt_radiance_map_update_distributions(self.h))

    def radiance_map_download(self):
        nv = self.n_vol
        d = dict(q=np.zeros((nv, CELLS), np.float32), cdf=np.zeros((nv, CELLS), np.float32), visits=np.zeros((nv, CELLS), np.uint32),
                 irradiance=np.zeros(nv, np.float32), pos=np.zeros((nv, 3), np.float32), nrm=np.zeros((nv, 3), np.float32), surface=np.zeros(nv, np.int32))
        self._ck(self.L.rlpt_radiance_map_download(self.h, *[_p(d[k]) for k in ("q", "cdf", "visits", "irradiance", "pos", "nrm", "surface")]))
        return d

    def radiance_map_delta(self):
        nv = self.n_vol
        s, c = np.zeros((nv, CELLS), np.float32), np.zeros((nv, CELLS), np.uint32)
        self._ck(self.L.rlpt_radiance_map_delta_download(self.h, _p(s), _p(c)))
        return s, c

    def radiance_map_save_q(self, path):
        self._ck(self.L.rlpt_radiance_map_save_q(self.h, path.encode()))

    def radiance_map_load_q(self, path):
        self._ck(self.L.rlpt_radiance_map_load_q(self.h, path.encode()))

    # ---- rendering
    def render_default(self, frames=1):
        self._ck(self.L.rlpt_render_default(self.h, int(frames)))

    def render_sarsa(self, frames=1):
        self._ck(self.L.rlpt_render_sarsa(self.h, int(frames)))

    def render_pretrained(self, frames=1):
        self._ck(self.L.rlpt_render_pretrained(self.h, int(frames)))

    def render_neuralq(self, frames=1, batch=4096):
        self._ck(self.L.rlpt_render_neuralq(self.h, int(frames), int(batch)))
        loss = ctypes.c_double()
        self._ck(self.L.rlpt_neuralq_last_loss(self.h, ctypes.byref(loss)))
        return loss.value

    def neuralq_set_hyper(self, learning_rate=1e-3, epsilon_start=0.05, epsilon_decay=0.01, epsilon_min=0.05):
        self._ck(self.L.rlpt_neuralq_set_hyper(self.h, ctypes.c_float(learning_rate), ctypes.c_float(epsilon_start), ctypes.c_float(epsilon_decay), ctypes.c_float(epsilon_min)))

    def radiance_map_build_seconds(self):
        v = (ctypes.c_double * 4)()
        self._ck(self.L.rlpt_radiance_map_build_seconds(self.h, v))
        return dict(volumes_and_kd_tree_host=v[0], candidate_cells_host=v[1], upload_and_first_cdf_device=v[2], total=v[3])

    def set_max_direction(self, on):
        self._ck(self.L.rlpt_set_max_direction(self.h, int(bool(on))))

    def render_voronoi(self):
        self._ck(self.L.rlpt_render_voronoi(self.h))

    def render_sarsa_frozen(self, frames=1):
        self._ck(self.L.rlpt_render_sarsa_frozen(self.h, int(frames)))

    def sarsa_trace(self):
        self._ck(self.L.rlpt_sarsa_trace(self.h))

    def sarsa_merge(self):
        self._ck(self.L.rlpt_sarsa_merge(self.h))

    def frame_reset(self):
        self._ck(self.L.rlpt_frame_reset(self.h))

    def frame_allreduce(self):
        self._ck(self.L.rlpt_frame_allreduce(self.h))

    def frame_download(self, out=None):
        n = self.cfg.width * self.cfg.height
        if out is None:
            out = np.zeros((n, 3), np.float32)
        self._ck(self.L.rlpt_frame_download(self.h, _p(out)))
        return out

    def frame_download_argb(self):
        a = np.zeros((self.cfg.height, self.cfg.width), np.uint32)
        self._ck(self.L.rlpt_frame_download_argb(self.h, _p(a)))
        return a

    def frame_save_bmp(self, path):
        self._ck(self.L.rlpt_frame_save_bmp(self.h, path.encode()))

    def stats(self):
        s = Stats()
        self._ck(self.L.rlpt_stats(self.h, ctypes.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def stats_reset(self):
        self._ck(self.L.rlpt_stats_reset(self.h))

    def measure_fp32_peak(self):
        v = ctypes.c_double()
        self._ck(self.L.rlpt_measure_fp32_peak(self.h, ctypes.byref(v)))
        return v.value

    # ---- Neural-Q network
    def dqn_set_vertices(self, vertices):
        v = _f32(vertices).ravel()
        self._ck(self.L.rlpt_dqn_set_vertices(self.h, _p(v), len(v)))

    def dqn_init(self, seed=1984):
        self._ck(self.L.rlpt_dqn_init(self.h, ctypes.c_uint32(seed)))

    def dqn_load_text(self, path):
        self._ck(self.L.rlpt_dqn_load_text(self.h, path.encode()))

    def dqn_save_text(self, path):
        self._ck(self.L.rlpt_dqn_save_text(self.h, path.encode()))

    def dqn_param_count(self):
        n, k = ctypes.c_int(), ctypes.c_int()
        self._ck(self.L.rlpt_dqn_param_count(self.h, ctypes.byref(n), ctypes.byref(k)))
        return n.value, k.value

    def dqn_set_params(self, params):
        p = _f32(params).ravel()
        self._ck(self.L.rlpt_dqn_set_params(self.h, _p(p), len(p)))

    def dqn_get_params(self):
        n, _ = self.dqn_param_count()
        p = np.zeros(n, np.float32)
        self._ck(self.L.rlpt_dqn_get_params(self.h, _p(p), n))
        return p

    def dqn_forward(self, pos):
        pos = _f32(pos).reshape(-1, 3)
        q = np.zeros((len(pos), CELLS), np.float32)
        self._ck(self.L.rlpt_dqn_forward(self.h, _p(pos), len(pos), _p(q)))
        return q

    def dqn_train_batch(self, pos, actions, targets, apply_update=True):
        pos = _f32(pos).reshape(-1, 3)
        a = np.ascontiguousarray(actions, dtype=np.uint32); t = _f32(targets)
        loss = ctypes.c_float()
        self._ck(self.L.rlpt_dqn_train_batch(self.h, _p(pos), _p(a), _p(t), len(pos), int(bool(apply_update)), ctypes.byref(loss)))
        return loss.value

    def dqn_train_supervised(self, pos, targets144, apply_update=True):
        pos = _f32(pos).reshape(-1, 3); t = _f32(targets144).reshape(len(pos), 144)
        loss = ctypes.c_float()
        self._ck(self.L.rlpt_dqn_train_supervised(self.h, _p(pos), _p(t), len(pos), int(apply_update), ctypes.byref(loss)))
        return float(loss.value)

    def dqn_get_grads(self):
        n, _ = self.dqn_param_count()
        g = np.zeros(n, np.float32)
        self._ck(self.L.rlpt_dqn_get_grads(self.h, _p(g), n))
        return g

    def capture_rays(self, method, bounce, max_rays):
        org, dir = np.zeros((max_rays, 3), np.float32), np.zeros((max_rays, 3), np.float32)
        n = ctypes.c_int()
        self._ck(self.L.rlpt_capture_rays(self.h, int(method), int(bounce), _p(org), _p(dir), int(max_rays), ctypes.byref(n)))
        return org[:n.value].copy(), dir[:n.value].copy()


def image_from_frame(rgb, width, height):
    """frame buffer (pixel = x*height + y) -> (height, width, 3) image array"""
    return np.asarray(rgb, np.float32).reshape(width, height, 3).transpose(1, 0, 2)
