"""oracle/checkers.py -- TEST INFRASTRUCTURE ONLY.

ctypes front ends for the two checkers:
  * ``Oracle``     oracle/_build/liboracle.so  (oracle/rlpt_oracle.cpp, the CPU restatement)
  * ``Reference``  oracle/_ref/libref_host.so or libref_cuda.so (the unmodified reference engine behind
                   oracle/ref_harness.cu)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
A = 144


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE], env={**os.environ, "CXX": "g++"})
    return os.path.join(HERE, "_build", "liboracle.so")


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.L = ctypes.CDLL(path)
        self.L.orc_tri_area.restype = ctypes.c_float
        self.ns = self.nl = 0

    def threads(self):
        return self.L.orc_threads()

    def set_max_direction(self, on):
        self.L.orc_set_max_direction(int(bool(on)))

    def philox(self, seed, pixel, sample, bounce, purpose):
        u = np.zeros(4, np.float32)
        self.L.orc_philox(ctypes.c_uint32(seed), ctypes.c_uint32(pixel), ctypes.c_uint32(sample), ctypes.c_uint32(bounce), ctypes.c_uint32(purpose), _p(u))
        return u

    def scene_set(self, sv, srgb, lv, lrgb):
        sv, srgb, lv, lrgb = _f32(sv).reshape(-1, 9), _f32(srgb).reshape(-1, 3), _f32(lv).reshape(-1, 9), _f32(lrgb).reshape(-1, 3)
        self.ns, self.nl = len(sv), len(lv)
        assert self.L.orc_scene_set(_p(sv), _p(srgb), self.ns, _p(lv), _p(lrgb), self.nl) == 0

    def scene_normals(self):
        sn, sl = np.zeros((self.ns, 3), np.float32), np.zeros(self.ns, np.float32)
        ln, ll = np.zeros((self.nl, 3), np.float32), np.zeros(self.nl, np.float32)
        self.L.orc_scene_normals(_p(sn), _p(sl), _p(ln), _p(ll))
        return sn, sl, ln, ll

    def closest_hit(self, org, dir, screen_height, fma_mode):
        org, dir = _f32(org).reshape(-1, 3), _f32(dir).reshape(-1, 3)
        n = len(org)
        ty, ix = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t, pos = np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        self.L.orc_closest_hit(_p(org), _p(dir), n, int(screen_height), int(fma_mode), _p(ty), _p(ix), _p(t), _p(pos))
        return ty, ix, t, pos

    def map(self, x, y):
        out = np.zeros(3, np.float32)
        self.L.orc_map(ctypes.c_float(x), ctypes.c_float(y), _p(out))
        return out

    def grid_dir(self, gx, gy, pos, nrm):
        gx, gy, pos, nrm = _f32(gx), _f32(gy), _f32(pos), _f32(nrm)
        out = np.zeros((len(gx), 3), np.float32)
        self.L.orc_grid_dir(_p(gx), _p(gy), len(gx), _p(pos), _p(nrm), _p(out))
        return out

    def uniform_hemisphere(self, nrm, r1, r2):
        out = np.zeros(3, np.float32)
        self.L.orc_uniform_hemisphere(_p(_f32(nrm)), ctypes.c_float(r1), ctypes.c_float(r2), _p(out))
        return out

    def tri_area(self, i):
        return self.L.orc_tri_area(int(i))

    def rmap_build(self, area_per_sample=0.001):
        self.nv = self.L.orc_rmap_build(ctypes.c_float(area_per_sample))
        return self.nv

    def rmap_counts(self):
        a, b = ctypes.c_int(), ctypes.c_int()
        self.L.orc_rmap_counts(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def rmap_volumes(self):
        nv, _ = self.rmap_counts()
        pos, nrm, surf = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros(nv, np.int32)
        self.L.orc_rmap_get_volumes(_p(pos), _p(nrm), _p(surf))
        return pos, nrm, surf

    def rmap_tree(self):
        _, nt = self.rmap_counts()
        dim, leaf = np.zeros(nt, np.int32), np.zeros(nt, np.int32)
        left, right = np.zeros(nt, np.uint32), np.zeros(nt, np.uint32)
        data, pos, nrm = np.zeros(nt, np.float32), np.zeros((nt, 3), np.float32), np.zeros((nt, 3), np.float32)
        self.L.orc_rmap_get_tree(_p(dim), _p(leaf), _p(left), _p(right), _p(data), _p(pos), _p(nrm))
        return dict(dim=dim, leaf=leaf, left=left, right=right, data=data, pos=pos, nrm=nrm)

    def rmap_state(self):
        nv, _ = self.rmap_counts()
        q, cdf = np.zeros((nv, A), np.float32), np.zeros((nv, A), np.float32)
        vis, irr = np.zeros((nv, A), np.uint32), np.zeros(nv, np.float32)
        self.L.orc_rmap_get_state(_p(q), _p(cdf), _p(vis), _p(irr))
        return q, cdf, vis, irr

    def rmap_acc(self):
        nv, _ = self.rmap_counts()
        s, c = np.zeros((nv, A), np.float64), np.zeros((nv, A), np.uint32)
        self.L.orc_rmap_get_acc(_p(s), _p(c))
        return s, c

    def rmap_set_acc(self, s, c):
        s, c = np.ascontiguousarray(s, dtype=np.float64), np.ascontiguousarray(c, dtype=np.uint32)
        self.L.orc_rmap_set_acc(_p(s), _p(c))

    def rmap_set_q(self, q):
        q = _f32(q)
        self.L.orc_rmap_set_q(_p(q))

    def rmap_update_distributions(self):
        self.L.orc_rmap_update_distributions()

    def rmap_merge_frame(self):
        self.L.orc_rmap_merge_frame()

    def find_closest(self, pos, nrm, fma_mode, max_dist=0.003):
        pos, nrm = _f32(pos).reshape(-1, 3), _f32(nrm).reshape(-1, 3)
        out = np.zeros(len(pos), np.int32)
        self.L.orc_find_closest(_p(pos), _p(nrm), len(pos), ctypes.c_float(max_dist), int(fma_mode), _p(out))
        return out

    def sample_sector(self, cdf_row, r):
        pdf = ctypes.c_float()
        s = self.L.orc_sample_sector(_p(_f32(cdf_row)), ctypes.c_float(r), ctypes.byref(pdf))
        return s, pdf.value

    def sample_sector_clamped(self, cdf_row, r):
        pdf = ctypes.c_float()
        s = self.L.orc_sample_sector_clamped(_p(_f32(cdf_row)), ctypes.c_float(r), ctypes.byref(pdf))
        return s, pdf.value

    def cell_cos(self, vol):
        out = np.zeros(A, np.float32)
        self.L.orc_cell_cos(int(vol), _p(out))
        return out

    def render_frame(self, method, width, height, spp, sample0=0, max_bounces=80, env=0.0, seed=1984, cam=(0, 0, -3),
                     yaw_y=0.0, yaw_x=0.0, fma_mode=1, td_mode=1, clamp_last_bin=1, max_dist=0.003):
        out = np.zeros((width * height, 3), np.float32)
        stats = np.zeros(4, np.float64)
        cam = _f32(cam)
        self.L.orc_render_frame(int(method), int(width), int(height), int(spp), int(sample0), int(max_bounces), ctypes.c_float(env),
                                ctypes.c_uint(seed), _p(cam), ctypes.c_float(yaw_y), ctypes.c_float(yaw_x), int(fma_mode), int(td_mode),
                                int(clamp_last_bin), ctypes.c_float(max_dist), _p(out), _p(stats))
        return out, dict(total_path_length=stats[0], zero_contribution=stats[1], failed=stats[2], paths=stats[3])


class Reference:
    """The unmodified reference engine. kind: 'host' (g++ shim build) or 'cuda' (its own kernels, GPU box only)."""

    def __init__(self, kind="host", suffix=""):
        path = os.path.join(HERE, "_ref", "libref_%s%s.so" % (kind, suffix))
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run oracle/build_ref.sh where /root/reference exists)")
        self.kind = kind
        self.L = ctypes.CDLL(path)
        w, h, s, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.L.ref_dims(ctypes.byref(w), ctypes.byref(h), ctypes.byref(s), ctypes.byref(b))
        self.width, self.height, self.spp, self.max_bounces = w.value, h.value, s.value, b.value

    @staticmethod
    def available(kind="host", suffix=""):
        return os.path.exists(os.path.join(HERE, "_ref", "libref_%s%s.so" % (kind, suffix)))

    def threads(self):
        return self.L.ref_threads()

    def scene_cornell(self):
        assert self.L.ref_scene_cornell() == 0

    def scene_obj(self, path, lights_in_obj):
        rc = self.L.ref_scene_obj(path.encode(), int(bool(lights_in_obj)))
        assert rc == 0, rc

    def scene_arrays(self, sv, srgb, lv, lrgb):
        sv, srgb, lv, lrgb = _f32(sv).reshape(-1, 9), _f32(srgb).reshape(-1, 3), _f32(lv).reshape(-1, 9), _f32(lrgb).reshape(-1, 3)
        assert self.L.ref_scene_arrays(_p(sv), _p(srgb), len(sv), _p(lv), _p(lrgb), len(lv)) == 0

    def scene_get(self):
        ns, nl = ctypes.c_int(), ctypes.c_int()
        assert self.L.ref_scene_counts(ctypes.byref(ns), ctypes.byref(nl)) == 0
        ns, nl = ns.value, nl.value
        d = dict(sv=np.zeros((ns, 9), np.float32), srgb=np.zeros((ns, 3), np.float32), snrm=np.zeros((ns, 3), np.float32), slum=np.zeros(ns, np.float32),
                 lv=np.zeros((nl, 9), np.float32), lrgb=np.zeros((nl, 3), np.float32), lnrm=np.zeros((nl, 3), np.float32), llum=np.zeros(nl, np.float32))
        assert self.L.ref_scene_get(*[_p(d[k]) for k in ("sv", "srgb", "snrm", "slum", "lv", "lrgb", "lnrm", "llum")]) == 0
        return d

    def camera(self, x, y, z, yaw_y=0.0, yaw_x=0.0):
        assert self.L.ref_camera(*[ctypes.c_float(v) for v in (x, y, z, yaw_y, yaw_x)]) == 0

    def closest_hit(self, org, dir):
        org, dir = _f32(org).reshape(-1, 3), _f32(dir).reshape(-1, 3)
        n = len(org)
        ty, ix = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t, pos = np.zeros(n, np.float32), np.zeros((n, 3), np.float32)
        assert self.L.ref_closest_hit(_p(org), _p(dir), n, _p(ty), _p(ix), _p(t), _p(pos)) == 0
        return ty, ix, t, pos

    def rmap_build(self):
        nv = self.L.ref_rmap_build()
        assert nv >= 0
        return nv

    def rmap_counts(self):
        a, b = ctypes.c_int(), ctypes.c_int()
        assert self.L.ref_rmap_counts(ctypes.byref(a), ctypes.byref(b)) == 0
        return a.value, b.value

    def rmap_volumes(self):
        nv, _ = self.rmap_counts()
        pos, nrm, surf = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros(nv, np.int32)
        self.L.ref_rmap_get_volumes(_p(pos), _p(nrm), _p(surf))
        return pos, nrm, surf

    def rmap_tree(self):
        _, nt = self.rmap_counts()
        dim, leaf = np.zeros(nt, np.int32), np.zeros(nt, np.int32)
        left, right = np.zeros(nt, np.uint32), np.zeros(nt, np.uint32)
        data, pos, nrm = np.zeros(nt, np.float32), np.zeros((nt, 3), np.float32), np.zeros((nt, 3), np.float32)
        self.L.ref_rmap_get_tree(_p(dim), _p(leaf), _p(left), _p(right), _p(data), _p(pos), _p(nrm))
        return dict(dim=dim, leaf=leaf, left=left, right=right, data=data, pos=pos, nrm=nrm)

    def rmap_state(self):
        nv, _ = self.rmap_counts()
        q, cdf = np.zeros((nv, A), np.float32), np.zeros((nv, A), np.float32)
        vis, irr = np.zeros((nv, A), np.uint32), np.zeros(nv, np.float32)
        self.L.ref_rmap_get_state(_p(q), _p(cdf), _p(vis), _p(irr))
        return q, cdf, vis, irr

    def rmap_set_q(self, q):
        q = _f32(q)
        assert self.L.ref_rmap_set_q(_p(q)) == 0

    def rmap_update_distributions(self):
        assert self.L.ref_rmap_update_distributions() == 0

    def find_closest(self, pos, nrm, on_device=False):
        pos, nrm = _f32(pos).reshape(-1, 3), _f32(nrm).reshape(-1, 3)
        out = np.zeros(len(pos), np.int32)
        assert self.L.ref_find_closest(_p(pos), _p(nrm), len(pos), _p(out), int(on_device)) == 0
        return out

    def grid_dir(self, vol, gx, gy):
        gx, gy = _f32(gx), _f32(gy)
        out = np.zeros((len(gx), 3), np.float32)
        assert self.L.ref_grid_dir(int(vol), _p(gx), _p(gy), len(gx), _p(out)) == 0
        return out

    def render_default(self, frames):
        out = np.zeros((self.width * self.height, 3), np.float32)
        stats = np.zeros((frames, 2), np.float64)
        assert self.L.ref_render_default(int(frames), _p(out), _p(stats)) == 0
        return out, stats

    def render_sarsa(self, frames, skip_frames=0):
        mean = np.zeros((self.width * self.height, 3), np.float32)
        last = np.zeros((self.width * self.height, 3), np.float32)
        stats = np.zeros((frames, 5), np.float64)
        assert self.L.ref_render_sarsa(int(frames), int(skip_frames), _p(mean), _p(last), _p(stats)) == 0
        return mean, last, stats


def philox_raw(oracle, ctr4, key2):
    c, k, o = np.array(ctr4, np.uint32), np.array(key2, np.uint32), np.zeros(4, np.uint32)
    oracle.L.orc_philox_raw(_p(c), _p(k), _p(o))
    return tuple(int(x) for x in o)


def mape_score(gt_rgb8, pred_rgb8):
    """Graphing/mape.py:10-21 of the reference: sum(|gt/255 - p/255| / ((gt + 0.01)/255)) / (H*W*3) on 8-bit RGB."""
    gt = np.asarray(gt_rgb8, np.float64)
    p = np.asarray(pred_rgb8, np.float64)
    return float(np.sum(np.abs(gt / 255.0 - p / 255.0) / ((gt + 0.01) / 255.0)) / gt.size)


def to_rgb8(rgb_float):
    """SDLScreen::PutPixelSDL colour conversion (G/sdl/sdl_screen.cpp:96-108): uint32(clamp(255*c, 0, 255)), truncation."""
    c = np.nan_to_num(np.asarray(rgb_float, np.float32), nan=0.0)
    return np.clip(np.float32(255.0) * c, 0.0, 255.0).astype(np.uint32).astype(np.uint8)


# ---------------------------------------------------------------------------------------------- Neural-Q network (numpy restatement)
DQN_ROWS = (200, 300, 200, 144)


def dqn_shapes(k_in):
    cols = (k_in, 200, 300, 200)
    return [(DQN_ROWS[l], cols[l]) for l in range(4)]


def dqn_split(params, k_in):
    """flat parameter vector (W1 b1 W2 b2 W3 b3 W4 b4, W row-major [out][in]) -> [(W, b)] * 4"""
    out, o = [], 0
    for r, c in dqn_shapes(k_in):
        W = np.asarray(params[o:o + r * c], np.float32).reshape(r, c); o += r * c
        b = np.asarray(params[o:o + r], np.float32); o += r
        out.append((W, b))
    assert o == len(params)
    return out


def dqn_forward_numpy(params, vertices, pos):
    """N/dq_network.cu:37-49 + N/fc_layer.cu:43-50 on the input of nn_rendering_helpers.cu:280-298:
    x_i = vertices[i] - pos[i % 3]; h = relu(b + W h) four times (ReLU on the output layer too), all float32."""
    vertices = np.asarray(vertices, np.float32).ravel(); pos = np.asarray(pos, np.float32).reshape(-1, 3)
    h = vertices[None, :] - np.tile(pos, (1, len(vertices) // 3))
    for W, b in dqn_split(params, len(vertices)):
        h = np.maximum(h @ W.T + b[None, :], np.float32(0)).astype(np.float32)
    return h


def dqn_forward_numpy_bf16(params, vertices, pos):
    """The same network function with the roundings of the tensor-core path (rlpt_dqn.cu): layer 1 in float32; the activations entering layers 2-4 and
    the weights of layers 2-4 rounded to bfloat16; products accumulated wide. Networks whose outputs are small differences of large terms (the committed
    door-room network: pre-activations up to 3e5 behind outputs of order 1) lose accuracy to those roundings -- this restatement shows how much."""
    vertices = np.asarray(vertices, np.float32).ravel(); pos = np.asarray(pos, np.float32).reshape(-1, 3)
    layers = dqn_split(params, len(vertices))
    h = vertices[None, :] - np.tile(pos, (1, len(vertices) // 3))
    W, b = layers[0]
    h = np.maximum(h @ W.T + b[None, :], np.float32(0)).astype(np.float32)
    for W, b in layers[1:]:
        h = np.maximum(bf16_round(h).astype(np.float64) @ bf16_round(W).T.astype(np.float64) + b[None, :], 0.0).astype(np.float32)
    return h


def dynet_text_load(path):
    """DyNet TextFileSaver dump -> flat parameter vector + k_in ("#Parameter# /_i {rows,cols} nbytes ZERO_GRAD", values column-major)"""
    blocks = []
    with open(path) as f:
        lines = f.read().split("\n")
    k_in = None
    for i in range(0, 16, 2):
        dims = [int(x) for x in lines[i][lines[i].index("{") + 1:lines[i].index("}")].split(",")]
        vals = np.array(lines[i + 1].split(), np.float32)
        if len(dims) == 2:
            blocks.append(vals.reshape(dims[1], dims[0]).T.copy().ravel())         # column-major -> row-major
            if i == 0:
                k_in = dims[1]
        else:
            blocks.append(vals)
    return np.concatenate(blocks).astype(np.float32), k_in


def dynet_text_save(path, params, k_in):
    with open(path, "w") as f:
        for l, (W, b) in enumerate(dqn_split(params, k_in)):
            f.write("#Parameter# /_%d {%d,%d} %d ZERO_GRAD\n" % (2 * l, W.shape[0], W.shape[1], W.size * 16 + 1))
            f.write("".join("%+.8e " % v for v in W.T.ravel()) + "\n")
            f.write("#Parameter# /_%d {%d} %d ZERO_GRAD\n" % (2 * l + 1, len(b), len(b) * 16 + 1))
            f.write("".join("%+.8e " % v for v in b) + "\n")


def bf16_round(a):
    """round-to-nearest-even to bfloat16, returned as float32 (what __float2bfloat16_rn does)"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def dqn_loss_and_grads_numpy(params, vertices, pos, actions, targets, bf16=False, all_outputs=False):
    """G/deep_learning/neural_q_pathtracer.cu:476-512 in float64 numpy: loss = sum_b (target_b - Q(s_b)[a_b])^2 and its
    gradient w.r.t. every parameter (flat, parameter order). ReLU follows every layer, the output included.
    bf16=True rounds exactly what the tensor-core path rounds (weights of layers 2-4, hidden activations, deltas, the query
    point in the layer-1 weight gradient) so that ReLU masks agree and the comparison can be tight."""
    r = bf16_round if bf16 else (lambda a: a)
    vertices = np.asarray(vertices, np.float64).ravel(); pos = np.asarray(pos, np.float64).reshape(-1, 3)
    layers = [(W.astype(np.float64), b.astype(np.float64)) for W, b in dqn_split(params, len(vertices))]
    x0 = vertices[None, :] - np.tile(pos, (1, len(vertices) // 3))
    hs = [x0]
    for l, (W, b) in enumerate(layers):
        Wl = W if l == 0 else r(W).astype(np.float64)
        pre = hs[-1] @ Wl.T + b[None, :]
        if l == 0 and bf16:
            pre = pre.astype(np.float32).astype(np.float64)
        h = np.maximum(pre, 0.0)
        hs.append(r(h).astype(np.float64) if l < 3 else h.astype(np.float32).astype(np.float64) if bf16 else h)
    q = hs[-1]; n = len(pos); idx = np.arange(n); t = np.asarray(targets, np.float64)
    W4 = layers[3][0]
    if all_outputs:
        # NN_Q_Value_Trainer/Source/main.cu:110-117: sum_batches squared_distance(targets [144], Q(s)); targets is [n][144]
        t = t.reshape(n, -1)
        loss = float(((t - q) ** 2).sum())
        G = 2.0 * (q - t) * (q > 0)
        gw4, gb4 = G.T @ hs[3], G.sum(0)
        d3 = r((G @ W4) * (hs[3] > 0)).astype(np.float64)
    else:
        a = np.asarray(actions, np.int64)
        qa = q[idx, a]
        loss = float(((t - qa) ** 2).sum())
        g = 2.0 * (qa - t) * (qa > 0)
        gw4 = np.zeros_like(W4); gb4 = np.zeros(W4.shape[0])
        np.add.at(gw4, a, g[:, None] * hs[3] * (hs[3] > 0)); np.add.at(gb4, a, g)
        d3 = r(g[:, None] * W4[a] * (hs[3] > 0)).astype(np.float64)
    d2 = r((d3 @ r(layers[2][0]).astype(np.float64)) * (hs[2] > 0)).astype(np.float64)
    d1 = r((d2 @ r(layers[1][0]).astype(np.float64)) * (hs[1] > 0)).astype(np.float64)
    gw3, gb3 = d3.T @ hs[2], d3.sum(0)
    gw2, gb2 = d2.T @ hs[1], d2.sum(0)
    if bf16:                                                         # rank-3 form with the query point rounded to bf16
        s1 = d1.sum(0); G = d1.T @ r(pos).astype(np.float64)
        gw1 = s1[:, None] * vertices[None, :] - np.tile(G, (1, len(vertices) // 3)); gb1 = s1
    else:
        gw1, gb1 = d1.T @ x0, d1.sum(0)
    grads = [(gw1, gb1), (gw2, gb2), (gw3, gb3), (gw4, gb4)]
    return loss, np.concatenate([np.concatenate([gw.ravel(), gb]) for gw, gb in grads]).astype(np.float32)


class AdamNumpy:
    """DyNet AdamTrainer defaults restated (lr 1e-3, beta 0.9 / 0.999, eps 1e-8, gradient clipping at norm 5). DyNet's source is
    not in the image and the reference pins no version: parity unpinned, this is the published algorithm."""

    def __init__(self, n, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, clip=5.0):
        self.m, self.v, self.t = np.zeros(n), np.zeros(n), 0
        self.lr, self.b1, self.b2, self.eps, self.clip = lr, b1, b2, eps, clip

    def step(self, params, grads):
        g = np.asarray(grads, np.float64); norm = np.sqrt((g * g).sum())
        if self.clip > 0 and norm > self.clip:
            g = g * (self.clip / norm)
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g; self.v = self.b2 * self.v + (1 - self.b2) * g * g
        step = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        return (np.asarray(params, np.float64) - step * self.m / (np.sqrt(self.v) + self.eps)).astype(np.float32)
