"""oracle/nq_tracer_oracle.py -- TEST INFRASTRUCTURE ONLY (checker, never the product).

numpy restatement of the reference's Neural-Q TRAINING tracer, NeuralQPathtracer::render_frame
(G/deep_learning/neural_q_pathtracer.cu:226-600), for one pass over the pixels, tracing the Philox paths the product traces:

  initialise_ray                        neural_q_pathtracer.cu:603-643     camera ray per pixel, throughput 1, discount 1
  per bounce
    sample_batch_ray_directions_epsilon_greedy   nn_rendering_helpers.cu:330-389 (-> importance_sample_direction :391-489)
    trace_ray                           neural_q_pathtracer.cu:646-752     closest hit, reward / discount / throughput by hit type
    compute_td_targets                  nn_rendering_helpers.cu:91-140     reward + discount * max_a Q(s', a) cos(theta_a)
    loss                                neural_q_pathtracer.cu:476-512     sum_batch (target - Q(s)[a])^2
    sample_random_scene_pos_for_terminated_rays  nn_rendering_helpers.cu:241-277

It follows the product's documented deviations from the reference (cell-centre cosine weights, paths that run
out of bounces contribute nothing, the re-seeding fixes) and its counter-based random numbers. The network itself is NOT restated here: the
caller passes `q_fn(positions [n,3]) -> Q [n,144]` (the tests pass the library's own forward, whose numerics tests/test_gpu_dqn.py pins against
numpy), so this file checks the TRACER: sampling, tracing, rewards, TD targets, re-seeding, frame-buffer accumulation.
The closest hit comes from the pinned CPU oracle (oracle/rlpt_oracle.cpp through checkers.Oracle)."""
import numpy as np

F = np.float32
GRID, CELLS = 12, 144
RHO = F(1.0) / (F(2.0) * F(3.1415926535))                  # G/constants/image_settings.h:13
GRID_RHO = F(1.0) / F(144.0)
RAY_EPS = F(0.00001)
PURPOSE_CAMERA, PURPOSE_NQ, PURPOSE_NQ_RESPAWN = 0, 2, 3


def philox4x32_10(seed, c0, c1, c2, c3):
    """Philox4x32-10, key (seed, 0), vectorised over the counter words (uint32 arrays). Returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & 0xFFFFFFFF for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(seed), np.uint64(0)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF), p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF); k1 = (k1 + np.uint64(0xBB67AE85)) & np.uint64(0xFFFFFFFF)
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def u01(x):                                                  # cuRAND's curand_uniform convention, (0, 1]
    return (x.astype(F) * F(2.3283064365386963e-10) + F(1.1641532182693481e-10)).astype(F)


def draw4(seed, pixel, sample, bounce, purpose):
    return [u01(c) for c in philox4x32_10(seed, pixel, np.full_like(pixel, sample), np.full_like(pixel, bounce), np.full_like(pixel, purpose))]


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def normalize_ref(v):                                        # glm::normalize as Ray::Ray applies it (G/rays/ray.cu:6-14)
    d = fma(v[:, 2], v[:, 2], fma(v[:, 0], v[:, 0], (v[:, 1] * v[:, 1]).astype(F)))
    inv = (F(1.0) / np.sqrt(d, dtype=F)).astype(F)
    return (v * inv[:, None]).astype(F)


def cell_centre_cos():
    k = np.arange(CELLS)
    a = F(2.0) * ((k // GRID).astype(F) + F(0.5)) * F(1.0 / 12.0) - F(1.0)
    b = F(2.0) * ((k % GRID).astype(F) + F(0.5)) * F(1.0 / 12.0) - F(1.0)
    r = np.maximum(np.abs(a), np.abs(b)).astype(F)
    return (F(1.0) - r * r).astype(F)


def tangent_frame(n):                                        # create_normal_coordinate_system (G/utils/hemisphere_helpers.cu:31-44)
    n = np.asarray(n, F)
    big = np.abs(n[:, 0]) > np.abs(n[:, 1])
    t = np.where(big[:, None], np.stack([n[:, 2], np.zeros_like(n[:, 0]), -n[:, 0]], 1), np.stack([np.zeros_like(n[:, 0]), -n[:, 2], n[:, 1]], 1)).astype(F)
    inv = (F(1.0) / np.sqrt((t[:, 0] * t[:, 0] + t[:, 1] * t[:, 1] + t[:, 2] * t[:, 2]).astype(F), dtype=F)).astype(F)
    T = (t * inv[:, None]).astype(F)
    B = np.stack([n[:, 1] * T[:, 2] - T[:, 1] * n[:, 2], n[:, 2] * T[:, 0] - T[:, 2] * n[:, 0], n[:, 0] * T[:, 1] - T[:, 0] * n[:, 1]], 1).astype(F)
    return T, B


def grid_to_direction(gx, gy, T, N, B):
    """convert_grid_pos_to_direction (G/utils/hemisphere_helpers.cu:96-105) with `map` (:134-226) in its four-quadrant concentric form:
    radius = max(|a|, |b|), angle by quadrant, y_h = 1 - r^2, sin(theta) = r sqrt(2 - r^2)"""
    a = (F(2.0) * (gx * F(1.0 / 12.0)).astype(F) - F(1.0)).astype(F); b = (F(2.0) * (gy * F(1.0 / 12.0)).astype(F) - F(1.0)).astype(F)
    q = F(0.78539816339744830962)
    upper = a > -b
    along_a = np.where(upper, a > b, a < b)
    den = np.where(along_a, a, b).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (np.where(along_a, b, a).astype(F) / den).astype(F)
    k = np.where(upper, np.where(along_a, F(0), F(2)), np.where(along_a, F(4), F(6))).astype(F)
    r = np.where(upper, den, -den).astype(F)
    phi = (q * (k + np.where(along_a, ratio, -ratio)).astype(F)).astype(F)
    phi = np.where(~upper & ~along_a & (b == 0), F(0), phi).astype(F)
    s, c = np.sin(phi, dtype=F), np.cos(phi, dtype=F)
    st = (r * np.sqrt((F(2.0) - r * r).astype(F), dtype=F)).astype(F)
    xh, yh, zh = (st * c).astype(F), (F(1.0) - r * r).astype(F), (st * s).astype(F)
    w = (T * xh[:, None] + N * yh[:, None] + B * zh[:, None]).astype(F)
    inv = (F(1.0) / np.sqrt((w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2]).astype(F), dtype=F)).astype(F)
    return (w * inv[:, None]).astype(F)


def camera_dir(px, py, u0, u1, width, height):               # Ray::sample_ray_through_pixel (G/rays/ray.cu:144-172), no rotation
    x = px.astype(F) + u0; y = py.astype(F) + u1
    v = np.stack([x - F(width) / F(2.0), y - F(height) / F(2.0), np.full_like(x, F(height))], 1).astype(F)
    return normalize_ref(v)


def sample_cells(q, cos, u0, u3, eps):
    """epsilon-greedy (nn_rendering_helpers.cu:330-389): explore with probability eps (uniform cell, pdf RHO), else sample the cell from
    Q(s, a) cos(theta_a) (:391-489). Sequential float32 running sums in cell order, as one thread of k_nqt_sample computes them."""
    n = len(u0)
    cell = np.zeros(n, np.int64); pdf = np.zeros(n, F)
    explore = u3 <= F(eps)
    cell[explore] = np.minimum((u0[explore] * F(CELLS)).astype(np.int64), CELLS - 1); pdf[explore] = RHO
    idx = np.nonzero(~explore)[0]
    if len(idx):
        qq = q[idx].astype(F)
        total = np.zeros(len(idx), F)
        for k in range(CELLS):
            total = fma(qq[:, k], cos[k], total)
        dead = ~(total > 0)
        wts = (qq * cos[None, :]).astype(F)
        if dead.any():
            wts[dead] = cos[None, :]
            t2 = np.zeros(int(dead.sum()), F)
            for k in range(CELLS):
                t2 = (t2 + cos[k]).astype(F)
            total[dead] = t2
        r = (u0[idx] * total).astype(F)
        run = np.zeros(len(idx), F); chosen = np.full(len(idx), -1, np.int64); wsel = np.zeros(len(idx), F)
        last = np.zeros(len(idx), np.int64); last_w = np.zeros(len(idx), F)
        for k in range(CELLS):
            w = wts[:, k]
            run = (run + w).astype(F)
            pos = w > 0
            take = (chosen < 0) & (run > r) & pos
            seen = pos & (chosen < 0)                            # `last` is tracked until the search stops (the stopping cell included)
            last[seen] = k; last_w[seen] = w[seen]
            chosen[take] = k; wsel[take] = w[take]
        none = chosen < 0
        chosen[none] = last[none]; wsel[none] = last_w[none]
        cell[idx] = chosen
        pdf[idx] = (RHO * ((wsel / total).astype(F) / GRID_RHO).astype(F)).astype(F)
    return cell, pdf


def nq_training_pass(orc, scene, q_fn, width, height, seed, sample_base, max_bounces, eps, env, cam, batch):
    """One pass of the training tracer over all pixels. Returns dict(accum [n,3] radiance sums, path_length_sum, zero_contribution, terminated,
    loss (sum over every batch of every bounce), steps, transitions per bounce)."""
    n = width * height
    sv = np.asarray(scene["sv"], F).reshape(-1, 3, 3); srgb = np.asarray(scene["srgb"], F).reshape(-1, 3)
    lrgb = np.asarray(scene["lrgb"], F).reshape(-1, 3)
    n_surf = len(sv)
    sn, slum, ln, llum = orc.scene_normals()
    T_s, B_s = tangent_frame(sn)
    brdf = (srgb / F(3.14159265358979323846)).astype(F)
    cos = cell_centre_cos()
    pix = np.arange(n, dtype=np.uint32)
    u = draw4(seed, pix, sample_base, 0, PURPOSE_CAMERA)
    d0 = camera_dir((pix // height).astype(np.int64), (pix % height).astype(np.int64), u[0], u[1], width, height)
    loc = np.tile(np.asarray(cam, F)[None, :], (n, 1)); gid = np.full(n, -1, np.int64)
    direction = d0.copy(); thr = np.ones((n, 3), F); state = np.zeros(n, np.int64)
    reward = np.zeros(n, F); discount = np.ones(n, F); action = np.zeros(n, np.int64)
    accum = np.zeros((n, 3), np.float64)
    st_len = st_zero = st_term = 0; loss = 0.0; steps = 0; transitions = []
    for b in range(max_bounces):
        sloc = loc.copy(); sgid = gid.copy()
        if b > 0:
            q = np.asarray(q_fn(loc), F)
            uu = draw4(seed, pix, sample_base, b, PURPOSE_NQ)
            ok = gid >= 0
            cell, pdf = sample_cells(q, cos, uu[0], uu[3], eps)
            g = np.maximum(gid, 0)
            nd = grid_to_direction((cell // GRID).astype(F) + uu[1], (cell % GRID).astype(F) + uu[2], T_s[g], sn[g], B_s[g])
            direction = np.where(ok[:, None], nd, direction).astype(F); action = np.where(ok, cell, action)
            scale = ((sn[g] * nd).sum(1, dtype=F) / pdf).astype(F)
            live = ok & (state == 0)
            thr = np.where(live[:, None], (thr * scale[:, None]).astype(F), thr)
        # trace_ray: every ray, alive or learning-only
        org = fma(np.full_like(direction, RAY_EPS), direction, loc)
        ty, ix, t, pos = orc.closest_hit(org, direction, height, 1)
        surf = ty == 2; light = ty == 1
        li = np.where(light, ix, 0); si = np.where(surf, ix, 0)                  # index into the light / surface tables (0 where not applicable)
        term = ~surf
        rgb_term = np.where(light[:, None], lrgb[li] if len(lrgb) else np.zeros((n, 3), F), np.full((n, 3), F(env), F)).astype(F)
        reward = np.where(light, (llum[li] if len(llum) else np.zeros(n, F)) * F(200.0), F(0)).astype(F)
        alive0 = state == 0
        add = term & alive0
        contrib = (rgb_term * thr).astype(F)
        nz = add & ((contrib != 0).any(1))
        accum[nz] += contrib[nz]
        st_len += int(add.sum()) * (b + 1); st_term += int(add.sum())
        st_zero += int((add & (contrib.sum(1, dtype=F) <= F(float.fromhex("0x1.3a92ap-12")))).sum())
        thr = np.where(add[:, None], contrib, thr)
        discount = np.where(term, F(0), slum[si]).astype(F)
        new_state = np.where(term, 1, state)
        out_of_bounces = surf & alive0 & (b + 1 >= max_bounces)
        st_len += int(out_of_bounces.sum()) * max_bounces; st_term += int(out_of_bounces.sum()); st_zero += int(out_of_bounces.sum())
        new_state = np.where(out_of_bounces, 2, new_state)
        cont = surf & alive0 & ~out_of_bounces
        thr = np.where(cont[:, None], (thr * brdf[si]).astype(F), thr)
        loc = np.where(surf[:, None], pos, loc).astype(F); gid = np.where(surf, ix, gid)
        state = new_state
        n_alive = int(cont.sum())
        if b > 0:
            for start in range(0, n, batch):
                sl = slice(start, min(start + batch, n))
                qn = np.asarray(q_fn(loc[sl]), F)
                best = np.maximum(F(0), (qn * cos[None, :]).astype(F).max(1)).astype(F)
                target = np.where(state[sl] != 1, fma(best, discount[sl], reward[sl]), reward[sl]).astype(F)
                qs = np.asarray(q_fn(sloc[sl]), F)
                qa = qs[np.arange(qs.shape[0]), action[sl]]
                loss += float(((qa.astype(np.float64) - target.astype(np.float64)) ** 2).sum()); steps += 1
                transitions.append((b, start, target.copy(), qa.copy()))
        # re-seed terminated rays on the geometry (they keep generating training data)
        dead = state == 1
        if dead.any():
            ur = draw4(seed, pix, sample_base, b, PURPOSE_NQ_RESPAWN)
            g2 = np.minimum((ur[0] * F(n_surf)).astype(np.int64), n_surf - 1)
            u1, u2 = ur[1].copy(), ur[2].copy()
            fold = (u1 + u2).astype(F) > F(1.0)
            u1[fold] = (F(1.0) - u1[fold]).astype(F); u2[fold] = (F(1.0) - u2[fold]).astype(F)
            v0 = sv[g2, 0]; e1 = (sv[g2, 1] - sv[g2, 0]).astype(F); e2 = (sv[g2, 2] - sv[g2, 0]).astype(F)
            p = fma(u2[:, None], e2, fma(u1[:, None], e1, v0))
            loc = np.where(dead[:, None], p, loc).astype(F); gid = np.where(dead, g2, gid); state = np.where(dead, 2, state)
        if n_alive == 0:
            break
    return dict(accum=accum, path_length_sum=st_len, zero_contribution=st_zero, terminated=st_term, loss=loss, steps=steps, transitions=transitions)
