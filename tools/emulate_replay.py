"""Lane-level NumPy emulation of the rotation-replay kernels' index logic (jacobi_vectors_kernel and the lean
variant): the round-robin slot ring, the orientation of a slot's two columns and the sign mask, checked against
the plain product V = J_1 J_2 ... of the recorded rotations.  CPU only; no CUDA code is executed."""
import numpy as np


def schedule(dd, step, k):
    per = dd - 1
    a0 = (step + k) % per
    b0 = (step + per - k) % per
    if k == 0:
        a0 = per
    return a0, b0


def reference(d, log):
    """V <- V J for every recorded rotation, with the solver's convention (p = min, q = max):
    col_p' = c col_p - s col_q, col_q' = s col_p + c col_q."""
    dd = d + (d & 1); np_ = dd // 2; per = dd - 1
    V = np.eye(dd)
    for g, row in enumerate(log):
        step = g % per
        for k in range(np_):
            a0, b0 = schedule(dd, step, k)
            p, q = min(a0, b0), max(a0, b0)
            c, s = row[k]
            vp, vq = V[:, p].copy(), V[:, q].copy()
            V[:, p] = c * vp - s * vq
            V[:, q] = s * vp + c * vq
    return V[:d, :d]


def lean(d, log):
    """What jacobi_vectors_lean_kernel does for every row: lane = slot, (va, vb) = the row's entries in the slot's
    columns, sign of s from the per-lane mask, ring shift by shuffles."""
    dd = d + (d & 1); np_ = dd // 2; per = dd - 1
    lanes = 32
    out = np.zeros((d, dd))
    flip = np.zeros((lanes, per), dtype=bool)
    for lane in range(lanes):
        for st in range(per):
            a0 = st + lane; b0 = st + per - lane
            if a0 >= per: a0 -= per
            if b0 >= per: b0 -= per
            if lane == 0: a0 = per
            flip[lane, st] = lane < np_ and a0 > b0
    for r in range(d):
        va = np.zeros(lanes); vb = np.zeros(lanes)
        for lane in range(np_):
            a0 = per if lane == 0 else lane
            b0 = 0 if lane == 0 else per - lane
            va[lane] = float(a0 == r); vb[lane] = float(b0 == r)
        st = 0
        for row in log:
            c = np.ones(lanes); s = np.zeros(lanes)
            for lane in range(lanes):
                cs = row[lane if lane < np_ else 0]
                c[lane], s[lane] = cs
            sg = np.where(flip[:, st], -s, s)
            na = c * va - sg * vb
            nb = sg * va + c * vb
            dn = np.concatenate([na[1:], na[-1:]])          # __shfl_down_sync(.., 1)
            up = np.concatenate([nb[:1], nb[:-1]])          # __shfl_up_sync(.., 1)
            if np_ > 1:
                nva = dn.copy(); nvb = up.copy()
                nva[0] = na[0]; nvb[0] = dn[0]
                nva[np_ - 1] = nb[np_ - 1]
                va, vb = nva, nvb
            else:
                va, vb = na, nb
            st = 0 if st + 1 == per else st + 1
        assert st == 0
        for lane in range(np_):
            a0 = per if lane == 0 else lane
            b0 = 0 if lane == 0 else per - lane
            out[r, a0] = va[lane]; out[r, b0] = vb[lane]
    return out[:, :d]


if __name__ == '__main__':
    rng = np.random.RandomState(0)
    for d in (1, 2, 3, 4, 7, 10, 31, 32, 33, 63, 64):
        dd = d + (d & 1); np_ = dd // 2; per = dd - 1
        sweeps = 2
        log = []
        for g in range(sweeps * per):
            row = []
            for k in range(np_):
                a0, b0 = schedule(dd, g % per, k)
                if max(a0, b0) >= d:                        # the bye of an odd dimension: identity
                    row.append((1.0, 0.0))
                else:
                    t = rng.uniform(-1, 1); c = 1 / np.sqrt(1 + t * t)
                    row.append((c, t * c))
            log.append(row)
        ref = reference(d, log)
        got = lean(d, log)
        err = np.max(np.abs(ref - got))
        print('d = %2d  max |V_ref - V_lean| = %.1e' % (d, err))
        assert err < 1e-14
    print('ok')
