"""oracle/numpy_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, independently written restatement of the reference's tracer advection
(src/oce_adv_tra_driver.F90:46-646, src/oce_adv_tra_hor.F90:64-834, src/oce_adv_tra_ver.F90:244-434,
:635-695, src/oce_adv_tra_fct.F90:72-512) in vectorised NumPy: whole-array masks and
``np.add.at`` scatters instead of the Fortran loops the C oracle (adv_oracle.c) transcribes.  It
exists to pin the C oracle: two restatements written in different styles
must agree to round-off.  Scatter order is kept serial (``np.add.at`` is unbuffered and processes
indices in order; node 1 and node 2 of an edge are interleaved), so the agreement is in practice
exact.

Covers: hor UPW1 / MUSCL / MFCT, ver UPW1 / QR4C / CDIFF / PPM, lim FCT / none, use_wsplit (the LO
vertical flux on w and the implicit Thomas sweep adv_tra_vert_impl on w_i, src/oce_adv_tra_ver.F90:90-240,
:438-631, src/oce_adv_tra_driver.F90:320-334), one rank.  Arrays are (column, level) = transposes of the Fortran shapes; levels are 1-based in
masks (``lev``).
"""
from __future__ import annotations

import numpy as np

R_EARTH = 6367500.0   # src/oce_modules.F90:29


class NumpyAdv:
    def __init__(self, mesh, state, nboundary_lay):
        m = self.m = mesh
        self.L, self.nl, self.N, self.Nh, self.E = m.L, m.nl, m.N, m.Nh, m.E
        f = lambda t: np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64)
        self.uv, self.w, self.we = f(state.uv), f(state.w), f(state.w_e)
        self.use_wsplit = bool(getattr(state, "use_wsplit", False))
        self.wi = f(state.w_i) if (self.use_wsplit and getattr(state, "w_i", None) is not None) else None
        self.helem, self.hnode, self.hnode_new = f(state.helem), f(state.hnode), f(state.hnode_new)
        self.zbar, self.Z = f(state.zbar_3d_n), f(state.Z_3d_n)
        self.nb = np.asarray(nboundary_lay)
        L = self.L
        self.lev = np.arange(1, L + 1)[None, :]                     # layer index nz
        self.n1 = m.edges[:, 0].astype(np.int64) - 1
        self.n2 = m.edges[:, 1].astype(np.int64) - 1
        self.el1 = m.edge_tri[:, 0].astype(np.int64) - 1
        self.has2 = m.edge_tri[:, 1] > 0
        self.el2 = np.where(self.has2, m.edge_tri[:, 1].astype(np.int64) - 1, 0)
        nu1 = m.ulevels[self.el1][:, None]; nl1 = (m.nlevels[self.el1] - 1)[:, None]
        nu2 = np.where(self.has2, m.ulevels[self.el2], 0)[:, None]
        nl2 = np.where(self.has2, m.nlevels[self.el2] - 1, 0)[:, None]
        nl12, nu12 = np.minimum(nl1, nl2), np.maximum(nu1, nu2)
        lev = self.lev
        # level ranges A-E of oce_adv_tra_hor.F90:127-160
        A = (lev >= nu1) & (lev <= nu12 - 1)
        B = (nu2 > 0) & (lev >= nu2) & (lev <= nu12 - 1)
        Cc = (lev >= nu12) & (lev <= nl12)
        D = (lev >= nl12 + 1) & (lev <= nl1)
        Ee = (lev >= nl12 + 1) & (lev <= nl2)
        self.use1, self.use2 = A | Cc | D, B | Cc | Ee
        # scatter range of oce_adv_tra_driver.F90:154-156
        lo = np.where(nu2 > 0, np.minimum(nu1, nu2), nu1)
        self.scat = (lev >= lo) & (lev <= np.maximum(nl1, nl2))
        uln = m.ulevels_nod2D[:, None]; nln = m.nlevels_nod2D[:, None]
        self.nvalid = (lev >= uln) & (lev <= nln - 1)               # (Nh, L) valid layers of a node
        self.uln, self.nln = uln, nln
        a = R_EARTH * m.elem_cos[self.el1]
        a = np.where(self.has2, 0.5 * (a + R_EARTH * m.elem_cos[self.el2]), a)
        self.cx = m.edge_dxdy[:, 0] * a
        self.cy = m.edge_dxdy[:, 1] * R_EARTH

    # ------------------------------------------------------------------ horizontal
    def volflux(self):
        c = self.m.edge_cross_dxdy
        u1, v1 = self.uv[self.el1, :, 0], self.uv[self.el1, :, 1]
        u2, v2 = self.uv[self.el2, :, 0], self.uv[self.el2, :, 1]
        f1 = (-v1 * c[:, 0:1] + u1 * c[:, 1:2]) * self.helem[self.el1]
        f2 = (v2 * c[:, 2:3] - u2 * c[:, 3:4]) * self.helem[self.el2]
        return np.where(self.use1 & self.use2, f1 + f2, np.where(self.use1, f1, np.where(self.use2, f2, 0.0)))

    def hor_upw1(self, ttf, Q, flux_in):
        t1, t2 = ttf[self.n1], ttf[self.n2]
        new = -0.5 * (t1 * (Q + np.abs(Q)) + t2 * (Q - np.abs(Q))) - flux_in
        return np.where(self.use1 | self.use2, new, flux_in)

    def hor_ho(self, kind, ttf, grad, Q, num_ord, flux_in):
        if kind == "UPW1":
            return self.hor_upw1(ttf, Q, flux_in)
        t1, t2 = ttf[self.n1], ttf[self.n2]
        cx, cy = self.cx[:, None], self.cy[:, None]
        d = 2.0 * (t2 - t1)
        Tm2 = (d + cx * grad[:, :, 1] + cy * grad[:, :, 3]) / 6.0
        Tm1 = (d + cx * grad[:, :, 0] + cy * grad[:, :, 2]) / 6.0
        if kind == "MUSCL":                                              # c_lo, oce_adv_tra_hor.F90:411-412
            Tm1 = Tm1 * np.where(self.nb[self.n1][:, None] - self.lev >= 0, 1.0, 0.0)
            Tm2 = Tm2 * np.where(self.nb[self.n2][:, None] - self.lev >= 0, 1.0, 0.0)
        Tmean1, Tmean2 = t1 + Tm1, t2 - Tm2
        cHO = (Q + np.abs(Q)) * Tmean1 + (Q - np.abs(Q)) * Tmean2
        new = -0.5 * (1.0 - num_ord) * cHO - Q * num_ord * 0.5 * (Tmean1 + Tmean2) - flux_in
        return np.where(self.use1 | self.use2, new, flux_in)

    def scatter_edges(self, acc, contrib):
        """acc(n1) += c, acc(n2) -= c per edge in ascending order (oce_adv_tra_driver.F90:142-201)."""
        c = np.where(self.scat, contrib, 0.0)
        idx = np.stack([self.n1, self.n2], 1).ravel()
        val = np.stack([c, -c], 1).reshape(-1, self.L)
        np.add.at(acc, idx, val)

    # ------------------------------------------------------------------ vertical (nl interfaces)
    def _iface(self):
        k = np.arange(1, self.nl + 1)[None, :]
        return k, self.uln[: self.N], self.nln[: self.N]

    def ver_upw1(self, w, ttf, flux_in):
        N, L, nl = self.N, self.L, self.nl
        k, nzmin, nzmax = self._iface()
        t = np.zeros((N, nl + 1)); t[:, 1:L + 1] = ttf[:N]             # t[:, k] = ttf(k), t[:,0] unused
        tk, tkm1 = t[:, 1:], t[:, :-1]
        W, A = w[:N], self.m.area[:N]
        out = flux_in.copy()
        top = k == nzmin
        out = np.where(top, -W * tk * A - out, out)
        out = np.where(k == nzmax, 0.0 - out, out)
        mid = (k >= nzmin + 1) & (k <= nzmax - 1)
        out = np.where(mid, -0.5 * (tk * (W + np.abs(W)) + tkm1 * (W - np.abs(W))) * A - out, out)
        return out

    def ver_cdiff(self, w, ttf, flux_in):
        N, L, nl = self.N, self.L, self.nl
        k, nzmin, nzmax = self._iface()
        t = np.zeros((N, nl + 1)); t[:, 1:L + 1] = ttf[:N]
        tk, tkm1 = t[:, 1:], t[:, :-1]
        W, A = w[:N], self.m.area[:N]
        tv = np.where(k == nzmin, -W * tk * A, -(0.5 * (tkm1 + tk)) * W * A)
        return np.where((k >= nzmin) & (k <= nzmax - 1), tv - flux_in, flux_in)

    def ver_qr4c(self, w, ttf, num_ord, flux_in):
        N, L, nl = self.N, self.L, self.nl
        k, nzmin, nzmax = self._iface()
        pad = lambda a: np.concatenate([np.zeros((N, 2)), a[:N, :L], np.zeros((N, 2))], 1)   # index k+1 -> level k
        t, Z = pad(ttf), pad(self.Z)
        g = lambda a, s: a[:, 2 + s: 2 + s + nl]                       # a(k+s) for k = 1..nl
        t0, tm1, tm2, tp1 = g(t, 0), g(t, -1), g(t, -2), g(t, 1)
        z0, zm1, zm2, zp1 = g(Z, 0), g(Z, -1), g(Z, -2), g(Z, 1)
        W, A, zb = w[:N], self.m.area[:N], self.zbar[:N]
        out = flux_in.copy()
        out = np.where(k == nzmin, -t0 * W * A - out, out)
        cen = -0.5 * (tm1 + t0) * W * A
        out = np.where(k == nzmin + 1, cen - out, out)
        out = np.where(k == nzmax - 1, cen - out, out)
        out = np.where(k == nzmax, 0.0 - out, out)
        inner = (k >= nzmin + 2) & (k <= nzmax - 2)
        with np.errstate(divide="ignore", invalid="ignore"):
            qc = (tm1 - t0) / (zm1 - z0)
            qu = (t0 - tp1) / (z0 - zp1)
            qd = (tm2 - tm1) / (zm2 - zm1)
            T1 = t0 + (2 * qc + qu) * (zb - z0) / 3.0
            T2 = tm1 + (2 * qc + qd) * (zb - zm1) / 3.0
            Tm = (W + np.abs(W)) * T1 + (W - np.abs(W)) * T2
            val = (-0.5 * (1.0 - num_ord) * Tm - num_ord * (0.5 * (T1 + T2)) * W) * A - out
        return np.where(inner, val, out)

    def ver_ppm(self, dt, w, ttf, flux_in):
        """adv_tra_vert_ppm (src/oce_adv_tra_ver.F90:487-627), whole-array form.  Arrays carry a guard of 2
        columns on each side: index k+2 <-> level / interface k (1-based)."""
        N, L, nl = self.N, self.L, self.nl
        nzmin, nzmax = self.uln[:N], self.nln[:N]
        G = 3
        pad = lambda a, n: np.concatenate([np.zeros((N, G)), a[:N, :n], np.zeros((N, G + nl - n))], 1)
        t, h, hold = pad(ttf, L), pad(self.hnode_new, L), pad(self.hnode, L)
        W, A = pad(w, nl), pad(self.m.area, nl)
        kk = np.arange(1 - G, nl + 1 + G)[None, :]                       # level index of every padded column
        sh = lambda a, s: np.roll(a, -s, axis=1)                          # a(k+s) at position k
        with np.errstate(divide="ignore", invalid="ignore"):
            # interface values tv(nz+1) of the loop :514-581 evaluated at every nz, masked afterwards
            dm1, dj, dp1, dp2 = sh(h, -1), h, sh(h, 1), sh(h, 2)
            tm1, t0, tp1, tp2 = sh(t, -1), t, sh(t, 1), sh(t, 2)
            dlt = dj / (dm1 + dj + dp1) * ((2.0 * dm1 + dj) / (dp1 + dj) * (tp1 - t0) + (dj + 2.0 * dp1) / (dm1 + dj) * (t0 - tm1))
            dlt1 = dp1 / (dj + dp1 + dp2) * ((2.0 * dj + dp1) / (dp2 + dp1) * (tp2 - tp1) + (dp1 + 2.0 * dp2) / (dj + dp1) * (tp1 - t0))
            sgn = lambda x: np.where(np.signbit(x), -1.0, 1.0)            # sign(1.0, x)
            lim = np.minimum(np.minimum(np.abs(dlt), 2.0 * np.abs(tp1 - t0)), 2.0 * np.abs(t0 - tm1)) * sgn(dlt)
            dlt = np.where((tp1 - t0) * (t0 - tm1) > 0.0, lim, 0.0)
            lim1 = np.minimum(np.minimum(np.abs(dlt1), 2.0 * np.abs(tp2 - tp1)), 2.0 * np.abs(tp1 - t0)) * sgn(dlt1)
            dlt1 = np.where((tp2 - tp1) * (tp1 - t0) > 0.0, lim1, 0.0)
            tvn = t0 + dj / (dj + dp1) * (tp1 - t0) + 1.0 / (dm1 + dj + dp1 + dp2) * (
                (2.0 * dp1 * dj) / (dj + dp1) * ((dm1 + dj) / (2.0 * dj + dp1) - (dp2 + dp1) / (2.0 * dp1 + dj)) * (tp1 - t0)
                - dj * (dm1 + dj) / (2.0 * dj + dp1) * dlt1 + dp1 * (dp1 + dp2) / (dj + 2.0 * dp1) * dlt)
        tv = np.zeros_like(t)
        inner = (kk >= nzmin + 1) & (kk <= nzmax - 3)                     # loop index nz, writes tv(nz+1)
        tv = np.where(sh(inner, -1), sh(tvn, -1), tv)
        # the four special interfaces in the order the reference assigns them (:497-507): later wins
        rows = np.arange(N)
        col = lambda k: (k + G - 1).ravel()                               # padded column of level k (1-based)
        tcol = lambda k: t[rows, col(k)]
        tv[rows, col(nzmin)] = tcol(nzmin)
        tv[rows, col(nzmin + 1)] = 0.5 * (tcol(nzmin) + tcol(nzmin + 1))
        tv[rows, col(nzmax - 1)] = 0.5 * (tcol(nzmax - 2) + tcol(nzmax - 1))
        tv[rows, col(nzmax)] = tcol(nzmax - 1)
        # parabola of every layer (:585-604)
        aL, aR = tv.copy(), sh(tv, 1).copy()
        flat = (aR - t) * (t - aL) <= 0.0
        aL = np.where(flat, t, aL); aR = np.where(flat, t, aR)
        c1 = (aR - aL) * (t - 0.5 * (aL + aR)) > (aR - aL) ** 2 / 6.0
        aL = np.where(c1, 3.0 * t - 2.0 * aR, aL)
        c2 = (aR - aL) * (t - 0.5 * (aR + aL)) < -((aR - aL) ** 2) / 6.0
        aR = np.where(c2, 3.0 * t - 2.0 * aL, aR)
        aj = 6.0 * (t - 0.5 * (aL + aR))
        layer = (kk >= nzmin) & (kk <= nzmax - 1)
        Wn = sh(W, 1)
        skip = (W <= 0.0) & (Wn >= 0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            xu = np.minimum(W * dt / hold, 1.0)
            fu = (-aL - 0.5 * xu * (aR - aL + (1.0 - 2.0 / 3.0 * xu) * aj)) * A * W          # tvert(nz), W(nz) > 0
            xd = np.minimum(-Wn * dt / hold, 1.0)
            fd = (-aR + 0.5 * xd * (aR - aL - (1.0 - 2.0 / 3.0 * xd) * aj)) * sh(A, 1) * Wn   # tvert(nz+1), W(nz+1) < 0
        tvert = np.zeros_like(t)
        tvert = np.where(layer & ~skip & (W > 0.0), fu, tvert)
        dn = layer & ~skip & (Wn < 0.0)
        tvert = np.where(sh(dn, -1), sh(fd, -1), tvert)
        tvert[rows, col(nzmin)] = -tv[rows, col(nzmin)] * W[rows, col(nzmin)] * A[rows, col(nzmin)]
        tvert[rows, col(nzmax)] = 0.0
        tvert = tvert[:, G:G + nl]
        k = np.arange(1, nl + 1)[None, :]
        return np.where((k >= nzmin) & (k <= nzmax), tvert - flux_in, flux_in)

    def vert_impl(self, dt, w, ttf):
        """adv_tra_vert_impl (src/oce_adv_tra_ver.F90:120-236): implicit upwind vertical advection, Thomas sweep
        written level by level over all owned columns at once.  Returns the updated (Nh, L) field."""
        N, L = self.N, self.L
        m = self.m
        out = np.array(ttf, dtype=np.float64, copy=True)
        nzmin = self.uln[:N, 0].astype(np.int64); nzmax = self.nln[:N, 0].astype(np.int64)
        W, A, AV, H, T = w[:N], m.area[:N], m.areasvol[:N], self.hnode_new[:N], out[:N]
        rows = np.arange(N)
        at = lambda a, k: a[rows, np.clip(k - 1, 0, a.shape[1] - 1)]      # a(k) per column, k 1-based arrays
        a_ = np.zeros((N, L + 2)); b_ = np.zeros((N, L + 2)); c_ = np.zeros((N, L + 2)); tr = np.zeros((N, L + 2))
        zinv = 1.0 * dt
        for nz in range(1, L + 1):
            kz = np.full(N, nz)
            first, mid, last = kz == nzmin, (kz >= nzmin + 1) & (kz <= nzmax - 2), kz == nzmax - 1
            act = first | mid | last
            if not act.any():
                continue
            with np.errstate(divide="ignore", invalid="ignore"):
                v1 = zinv * at(A, kz) / at(AV, kz)
                v2 = zinv * at(A, kz + 1) / at(AV, kz)
            wk, wk1, h = at(W, kz), at(W, kz + 1), at(H, kz)
            a = np.where(first, 0.0, np.minimum(0.0, wk) * v1)
            b = np.where(first, h + wk * v1, h + np.maximum(0.0, wk) * v1)
            b = np.where(last & ~first, b, b - np.minimum(0.0, wk1) * v2)
            c = np.where(last & ~first, 0.0, -np.maximum(0.0, wk1) * v2)
            tm1, t0, tp1 = at(T, kz - 1), at(T, kz), at(T, kz + 1)
            r_first = -(b - h) * t0 - c * tp1
            r_mid = -a * tm1 - (b - h) * t0 - c * tp1
            r_last = -a * tm1 - (b - h) * t0
            r = np.where(first, r_first, np.where(mid, r_mid, r_last))
            a_[:, nz] = np.where(act, a, 0.0); b_[:, nz] = np.where(act, b, 0.0)
            c_[:, nz] = np.where(act, c, 0.0); tr[:, nz] = np.where(act, r, 0.0)
        cp = np.zeros((N, L + 2)); tp = np.zeros((N, L + 2))
        for nz in range(1, L + 1):
            kz = np.full(N, nz)
            first = kz == nzmin
            rest = (kz >= nzmin + 1) & (kz <= nzmax - 1)
            with np.errstate(divide="ignore", invalid="ignore"):
                mm = b_[:, nz] - cp[:, nz - 1] * a_[:, nz]
                cp[:, nz] = np.where(first, c_[:, nz] / b_[:, nz], np.where(rest, c_[:, nz] / mm, 0.0))
                tp[:, nz] = np.where(first, tr[:, nz] / b_[:, nz], np.where(rest, (tr[:, nz] - tp[:, nz - 1] * a_[:, nz]) / mm, 0.0))
        x = np.zeros((N, L + 2))
        for nz in range(L, 0, -1):
            kz = np.full(N, nz)
            bottom = kz == nzmax - 1
            above = (kz >= nzmin) & (kz <= nzmax - 2)
            x[:, nz] = np.where(bottom, tp[:, nz], np.where(above, tp[:, nz] - cp[:, nz] * x[:, nz + 1], 0.0))
        valid = self.nvalid[:N]
        out[:N] = np.where(valid, T + x[:, 1:L + 1], T)
        return out

    def ver(self, kind, w, ttf, num_ord, flux_in, dt=None):
        if kind == "UPW1":
            return self.ver_upw1(w, ttf, flux_in)
        if kind == "QR4C":
            return self.ver_qr4c(w, ttf, num_ord, flux_in)
        if kind == "CDIFF":
            return self.ver_cdiff(w, ttf, flux_in)
        if kind == "PPM":
            return self.ver_ppm(dt, w, ttf, flux_in)
        raise ValueError(kind)

    # ------------------------------------------------------------------ FCT limiter
    def fct(self, dt, ttf, lo, adf_h, adf_v):
        m, N, L = self.m, self.N, self.L
        big = 1.0e3
        tmax, tmin = np.maximum(lo, ttf), np.minimum(lo, ttf)           # a1
        en = m.elem2D_nodes.astype(np.int64) - 1
        emask = (self.lev >= m.ulevels[:, None]) & (self.lev <= (m.nlevels - 1)[:, None])
        amax = np.where(emask, np.maximum.reduce([tmax[en[:, j]] for j in range(3)]), -big)   # a2 (AUX)
        amin = np.where(emask, np.minimum.reduce([tmin[en[:, j]] for j in range(3)]), big)
        tvmax = np.full((N, L), -np.inf); tvmin = np.full((N, L), np.inf)
        for j in range(m.nod_in_elem2D.shape[1]):                        # a3
            have = (j < m.nod_in_elem2D_num[:N])[:, None]
            el = np.where(have[:, 0], m.nod_in_elem2D[:N, j].astype(np.int64) - 1, 0)
            tvmax = np.where(have, np.maximum(tvmax, amax[el]), tvmax)
            tvmin = np.where(have, np.minimum(tvmin, amin[el]), tvmin)
        nzmin, nzmax = self.uln[:N], self.nln[:N]
        lev = self.lev
        sh = lambda a, s, fill: np.concatenate([np.full((N, 1), fill), a, np.full((N, 1), fill)], 1)[:, 1 + s:1 + s + L]
        inner = (lev >= nzmin + 1) & (lev <= nzmax - 2)
        vmax = np.where(inner, np.maximum.reduce([sh(tvmax, -1, -np.inf), tvmax, sh(tvmax, 1, -np.inf)]), tvmax)
        vmin = np.where(inner, np.minimum.reduce([sh(tvmin, -1, np.inf), tvmin, sh(tvmin, 1, np.inf)]), tvmin)
        valid = self.nvalid[:N]
        inc_max = np.where(valid, vmax - lo[:N], 0.0)
        inc_min = np.where(valid, vmin - lo[:N], 0.0)
        plus = np.zeros((self.Nh, L)); minus = np.zeros((self.Nh, L))
        plus[:N] = np.where(valid, np.maximum(0.0, adf_v[:, :L]) + np.maximum(0.0, -adf_v[:, 1:L + 1]), 0.0)
        minus[:N] = np.where(valid, np.minimum(0.0, adf_v[:, :L]) + np.minimum(0.0, -adf_v[:, 1:L + 1]), 0.0)
        idx = np.stack([self.n1, self.n2], 1).ravel()
        f = np.where(self.scat, adf_h, 0.0)
        np.add.at(plus, idx, np.stack([np.maximum(0.0, f), np.maximum(0.0, -f)], 1).reshape(-1, L))
        np.add.at(minus, idx, np.stack([np.minimum(0.0, f), np.minimum(0.0, -f)], 1).reshape(-1, L))
        av, hn = m.areasvol[:N, :L], self.hnode_new[:N]
        with np.errstate(divide="ignore", invalid="ignore"):
            Rp = np.where(valid, np.minimum(1.0, inc_max / (plus[:N] * dt / av / hn + 1e-16)), 0.0)
            Rm = np.where(valid, np.minimum(1.0, inc_min / (minus[:N] * dt / av / hn - 1e-16)), 0.0)
        # b3 vertical
        k = np.arange(1, self.nl + 1)[None, :]
        pk = np.concatenate([Rp, np.ones((N, 1))], 1); mk = np.concatenate([Rm, np.ones((N, 1))], 1)
        pa = np.concatenate([np.ones((N, 1)), Rp], 1); ma = np.concatenate([np.ones((N, 1)), Rm], 1)   # layer k-1
        pos = adf_v >= 0.0
        ae_top = np.where(pos, np.minimum(1.0, pk), np.minimum(1.0, mk))
        ae_mid = np.where(pos, np.minimum(np.minimum(1.0, ma), pk), np.minimum(np.minimum(1.0, pa), mk))
        ae = np.where(k == nzmin, ae_top, np.where((k >= nzmin + 1) & (k <= nzmax - 1), ae_mid, 1.0))
        adf_v = ae * adf_v
        return Rp, Rm, adf_v

    def limit_h(self, adf_h, Rp_all, Rm_all):
        p1, m1, p2, m2 = Rp_all[self.n1], Rm_all[self.n1], Rp_all[self.n2], Rm_all[self.n2]
        ae = np.where(adf_h >= 0.0, np.minimum(np.minimum(1.0, p1), m2), np.minimum(np.minimum(1.0, m1), p2))
        return np.where(self.scat, ae * adf_h, adf_h)

    # ------------------------------------------------------------------ driver
    def do_oce_adv_tra(self, dt, tr, dttf_h, dttf_v, diag=None):
        """one tracer; accumulates into dttf_h / dttf_v (Nh, L) like the reference.  ``diag``: a dict that receives the
        optional diagnostics of the call -- tra_advhoriz, tra_advvert (Nh, L) for ltra_diag (driver :221-229, :307-318,
        :464-488) and dvd_trflx_hor (E, L), dvd_trflx_ver (N, nl) for ldiag_DVD (:263-296, :395-458)"""
        m, N, L, nl = self.m, self.N, self.L, self.nl
        f = lambda t: np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64)
        ttf, ttfAB, grad = f(tr.values), f(tr.valuesAB), f(tr.edge_up_dn_grad)
        Q = self.volflux()
        fct = tr.tra_adv_lim.strip() == "FCT"
        zE, zV = np.zeros((self.E, L)), np.zeros((N, nl))
        valid = self.nvalid[:N]
        av = m.areasvol[:N, :L]
        if fct:
            flo_h = self.hor_upw1(ttf, Q, zE)
            lo = np.zeros((self.Nh, L))
            self.scatter_edges(lo, flo_h)
            flo_v = self.ver_upw1(self.we, ttf, zV)
            if diag is not None:
                hnn_all = self.hnode_new
                with np.errstate(divide="ignore", invalid="ignore"):
                    diag["tra_advhoriz"] = np.where(self.nvalid, lo * dt / m.areasvol[:, :L] / hnn_all, 0.0)
                    tv = np.zeros((self.Nh, L))
                    tv[:N] = np.where(valid, (flo_v[:, :L] - flo_v[:, 1:L + 1]) * dt / av / hnn_all[:N], 0.0)
                diag["tra_advvert"] = tv
                diag["dvd_trflx_hor"], diag["dvd_trflx_ver"] = flo_h.copy(), flo_v.copy()
            with np.errstate(divide="ignore", invalid="ignore"):
                lo_new = (ttf[:N] * self.hnode[:N] + (lo[:N] + (flo_v[:, :L] - flo_v[:, 1:L + 1])) * dt / av) / self.hnode_new[:N]
            lo = np.where(self.nvalid, 0.0, 0.0)
            lo[:N] = np.where(valid, lo_new, 0.0)
            if self.use_wsplit:                                          # driver :320-334
                lo = self.vert_impl(dt, self.wi, lo)
                flo_v = self.ver_upw1(self.w, ttf, zV)
            adf_h = self.hor_ho(tr.tra_adv_hor.strip(), ttfAB, grad, Q, tr.tra_adv_ph, flo_h)
            adf_v = self.ver(tr.tra_adv_ver.strip(), self.w, ttfAB, tr.tra_adv_pv, flo_v, dt)
            Rp, Rm, adf_v = self.fct(dt, ttf, lo, adf_h, adf_v)
            Rp_all = np.zeros((self.Nh, L)); Rm_all = np.zeros((self.Nh, L))
            Rp_all[:N], Rm_all[:N] = Rp, Rm
            adf_h = self.limit_h(adf_h, Rp_all, Rm_all)
            dttf_v[:N] = np.where(valid, dttf_v[:N] - ttf[:N] * self.hnode[:N] + lo[:N] * self.hnode_new[:N], dttf_v[:N])
            self.keep = dict(fct_LO=lo, fct_plus=Rp_all, fct_minus=Rm_all)
        else:
            adf_h = self.hor_ho(tr.tra_adv_hor.strip(), ttfAB, grad, Q, tr.tra_adv_ph, zE)
            adf_v = self.ver(tr.tra_adv_ver.strip(), self.we, ttfAB, tr.tra_adv_pv, zV, dt)
        with np.errstate(divide="ignore", invalid="ignore"):
            dttf_v[:N] = np.where(valid, dttf_v[:N] + (adf_v[:, :L] - adf_v[:, 1:L + 1]) * dt / av, dttf_v[:N])
        # U3: per edge, node 1 then node 2, addend (flux*dt)/areasvol(nz,node)
        c = np.where(self.scat, adf_h, 0.0) * dt
        idx = np.stack([self.n1, self.n2], 1).ravel()
        a1, a2 = m.areasvol[self.n1, :L], m.areasvol[self.n2, :L]
        with np.errstate(divide="ignore", invalid="ignore"):
            val = np.stack([np.where(self.scat, c / a1, 0.0), np.where(self.scat, -(c / a2), 0.0)], 1).reshape(-1, L)
        np.add.at(dttf_h, idx, val)
        if diag is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                th = np.where(self.nvalid, dttf_h / self.hnode_new, 0.0)
                tv = np.where(self.nvalid, dttf_v / self.hnode_new, 0.0)
            if fct:                                                       # LO part + (antidiffusive part incl. what was there)
                diag["tra_advhoriz"] = np.where(self.nvalid, diag["tra_advhoriz"] + th, 0.0)
                diag["tra_advvert"] = np.where(self.nvalid, diag["tra_advvert"] + tv, 0.0)
                diag["dvd_trflx_hor"] = diag["dvd_trflx_hor"] + adf_h
                diag["dvd_trflx_ver"] = diag["dvd_trflx_ver"] + adf_v
            else:
                diag["tra_advhoriz"], diag["tra_advvert"] = th, tv
                diag["dvd_trflx_hor"], diag["dvd_trflx_ver"] = adf_h.copy(), adf_v.copy()
        return dttf_h, dttf_v


def init_tracers_AB(values, valuesold, ab_order=2, epsilon=0.1):
    """src/oce_tracer_mod.F90:45-54, :97-122: returns (valuesAB, new valuesold).  ``valuesold`` has the shape
    (Nh, L, ab_order-1) (the reference's (ab_order-1, nl-1, Nh) read in C order)."""
    v = np.asarray(values, dtype=np.float64)
    o = np.asarray(valuesold, dtype=np.float64)
    if ab_order == 2:
        vab = -(0.5 + epsilon) * o[..., 0] + (1.5 + epsilon) * v
        new = v[..., None].copy()
    elif ab_order == 3:
        vab = (5.0 * o[..., 1] - 16.0 * o[..., 0] + 23.0 * v) / 12.0
        new = np.stack([v, o[..., 0]], axis=-1)
    else:
        raise ValueError("Adams-Bashfort tracer order must be 2 or 3")
    return vab, new


def tracer_gradient_elements(mesh, ttf):
    """src/oce_tracer_mod.F90:171-180, vectorised: ttf (Nh, L) -> tr_xy (T, L, 2); zero outside ulevels..nlevels-1."""
    en = np.asarray(mesh.elem2D_nodes, dtype=np.int64) - 1
    g = np.asarray(mesh.gradient_sca, dtype=np.float64)
    t = np.asarray(ttf, dtype=np.float64)
    lev = np.arange(1, mesh.L + 1)[None, :]
    wet = (lev >= np.asarray(mesh.ulevels)[:mesh.T, None]) & (lev <= np.asarray(mesh.nlevels)[:mesh.T, None] - 1)
    out = np.zeros((mesh.T, mesh.L, 2))
    for c in range(2):
        v = g[:, 3 * c + 0, None] * t[en[:, 0]] + g[:, 3 * c + 1, None] * t[en[:, 1]] + g[:, 3 * c + 2, None] * t[en[:, 2]]
        out[:, :, c] = np.where(wet, v, 0.0)
    return out


def fill_up_dn_grad(mesh, tr_xy, edge_up_dn_tri):
    """src/oce_muscl_adv.F90:378-522, vectorised over edges and levels; the node mean runs over the
    nod_in_elem2D slots in their order (k = 1 .. nod_in_elem2D_num), like the reference's inner loop."""
    L, E = mesh.L, mesh.E
    tr = np.asarray(tr_xy, dtype=np.float64)
    nlev_e, ulev_e = np.asarray(mesh.nlevels), np.asarray(mesh.ulevels)
    area = np.asarray(mesh.elem_area, dtype=np.float64)
    nie = np.asarray(mesh.nod_in_elem2D, dtype=np.int64) - 1
    num = np.asarray(mesh.nod_in_elem2D_num)
    Nn = nie.shape[0]
    lev = np.arange(1, L + 1)[None, :]
    tvol = np.zeros((Nn, L)); tx = np.zeros((Nn, L)); ty = np.zeros((Nn, L))
    for k in range(nie.shape[1]):
        el = np.where(k < num, nie[:, k], 0)
        ok = (k < num)[:, None] & ~((nlev_e[el][:, None] - 1 < lev) | (lev < ulev_e[el][:, None]))
        a = area[el][:, None]
        tvol = np.where(ok, tvol + a, tvol)
        tx = np.where(ok, tx + tr[el, :, 0] * a, tx)
        ty = np.where(ok, ty + tr[el, :, 1] * a, ty)
    with np.errstate(invalid="ignore", divide="ignore"):
        gx, gy = tx / tvol, ty / tvol
    ed = np.asarray(mesh.edges, dtype=np.int64) - 1
    tri = np.asarray(edge_up_dn_tri, dtype=np.int64)
    n1, n2 = ed[:, 0], ed[:, 1]
    both = ((tri[:, 0] != 0) & (tri[:, 1] != 0))[:, None]
    nmin, umax = np.asarray(mesh.nlevels_nod2D_min), np.asarray(mesh.ulevels_nod2D_max)
    nln, uln = np.asarray(mesh.nlevels_nod2D), np.asarray(mesh.ulevels_nod2D)
    nzmin = np.maximum(umax[n1], umax[n2])[:, None]
    nzmax = np.minimum(nmin[n1], nmin[n2])[:, None]
    shared = both & (lev >= nzmin) & (lev <= nzmax - 1)
    out = np.zeros((E, L, 4))
    up, dn = np.maximum(tri[:, 0] - 1, 0), np.maximum(tri[:, 1] - 1, 0)
    for node, t_el, cx, cy in ((n1, up, 0, 2), (n2, dn, 1, 3)):
        col = (lev >= uln[node][:, None]) & (lev <= nln[node][:, None] - 1)
        # :388-430 run from ulevels_nod2D(node) to nzmin-1, but :445-485 start at nzmax WITHOUT the node's upper bound: under
        # a cavity (nzmax < ulevels_nod2D(node)) the reference divides 0/0 there and stores the NaN (nobody reads it)
        mean = np.where(both, (col & (lev <= nzmin - 1)) | ((lev >= nzmax) & (lev <= nln[node][:, None] - 1)), col)
        out[:, :, cx] = np.where(shared, tr[t_el, :, 0], np.where(mean, gx[node], 0.0))
        out[:, :, cy] = np.where(shared, tr[t_el, :, 1], np.where(mean, gy[node], 0.0))
    return out


def vert_vel_ale_core(mesh, uv, helem):
    """src/oce_ale.F90:2164-2310 (linfs, no Fer_GM), vectorised over levels, edges in ascending order (the
    order of the reference's scatter): uv (T, L, 2), helem (T, L) -> Wvel (Nh, nl), owned nodes completed."""
    L, nl, Nh, N = mesh.L, mesh.nl, mesh.Nh, mesh.N
    ed = np.asarray(mesh.edges, dtype=np.int64) - 1
    et = np.asarray(mesh.edge_tri, dtype=np.int64) - 1
    cr = np.asarray(mesh.edge_cross_dxdy, dtype=np.float64)
    nlev, ulev = np.asarray(mesh.nlevels), np.asarray(mesh.ulevels)
    lev = np.arange(1, L + 1)
    W = np.zeros((Nh, nl))
    for e in range(mesh.E):
        for k, sgn in ((0, 1.0), (1, -1.0)):
            el = et[e, k]
            if el < 0:
                continue
            wet = (lev >= ulev[el]) & (lev <= nlev[el] - 1)
            c = (uv[el, :, 1] * cr[e, 2 * k] - uv[el, :, 0] * cr[e, 2 * k + 1]) * helem[el]
            c = np.where(wet, sgn * c, 0.0)
            W[ed[e, 0], :L] = np.where(wet, W[ed[e, 0], :L] + c, W[ed[e, 0], :L])
            W[ed[e, 1], :L] = np.where(wet, W[ed[e, 1], :L] - c, W[ed[e, 1], :L])
    nln, uln = np.asarray(mesh.nlevels_nod2D), np.asarray(mesh.ulevels_nod2D)
    area = np.asarray(mesh.area, dtype=np.float64)
    for n in range(N):
        a, b = uln[n], nln[n] - 1                            # nzmin, nzmax (1-based layers)
        for nz in range(b, a - 1, -1):
            W[n, nz - 1] = W[n, nz - 1] + W[n, nz]
        W[n, a - 1:b] = W[n, a - 1:b] / area[n, a - 1:b]
    return W


def compute_cflz_and_split(mesh, hnode_new, dt, W, use_wsplit, wsplit_maxcfl):
    """compute_CFLz (src/oce_ale.F90:2939-2952) and compute_Wvel_split (:3033-3047), whole-array: the loop's
    CFL_z(nz) = c2 of the layer above + c1 of the layer itself collapses to |W(nz) dt/h(nz-1)| + |W(nz) dt/h(nz)|
    (first interface: c1 only, interface below the last layer: c2 only).  W (Nh, nl) -> (CFL_z, W_e, W_i)."""
    L, nl, Nh = mesh.L, mesh.nl, mesh.Nh
    nln, uln = np.asarray(mesh.nlevels_nod2D)[:, None], np.asarray(mesh.ulevels_nod2D)[:, None]
    k = np.arange(1, nl + 1)[None, :]
    h = np.asarray(hnode_new, dtype=np.float64)
    lay = (k[:, :L] >= uln) & (k[:, :L] <= nln - 1)                       # wet layers
    with np.errstate(divide="ignore", invalid="ignore"):
        c1 = np.where(lay, np.abs(W[:, :L] * dt / h), 0.0)                 # own layer below interface nz
        c2 = np.where(lay, np.abs(W[:, 1:] * dt / h), 0.0)                 # layer above interface nz+1
    cfl = np.zeros((Nh, nl))
    cfl[:, 1:] = c2
    cfl[:, :L] = cfl[:, :L] + c1
    wet_i = (k >= uln) & (k <= nln)
    we = np.where(wet_i, W, 0.0)
    wi = np.zeros_like(W)
    if use_wsplit:
        big = wet_i & (cfl > wsplit_maxcfl)
        dd = np.maximum(cfl - wsplit_maxcfl, 0.0) / max(wsplit_maxcfl, 1.e-12)
        we = np.where(big, (1.0 / (1.0 + dd)) * W, we)
        wi = np.where(big, (dd / (1.0 + dd)) * W, wi)
    return cfl, we, wi


def vert_vel_ale_zstar(mesh, zbar_3d_n, hnode, hnode_new, dt, W, hbar, hbar_old, water_flux):
    """src/oce_ale.F90:2539-2603, whole-array: returns (W, hnode_new) with the zstar correction applied to the owned,
    cavity-free columns."""
    N, L, nl = mesh.N, mesh.L, mesh.nl
    W, hn = np.array(W, dtype=np.float64), np.array(hnode_new, dtype=np.float64)
    zb = np.asarray(zbar_3d_n, dtype=np.float64)[:N]
    uln = np.asarray(mesh.ulevels_nod2D)[:N]
    nzmax = np.asarray(mesh.nlevels_nod2D_min)[:N] - 1
    rows = np.arange(N)
    dd1 = zb[rows, nzmax - 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        dd = (np.asarray(hbar)[:N] - np.asarray(hbar_old)[:N]) / (zb[rows, uln - 1] - dd1)
        dddt = dd / dt
    k = np.arange(1, nl + 1)[None, :]
    sel = (uln == 1)[:, None] & (k >= uln[:, None]) & (k <= (nzmax - 1)[:, None])          # layers nzmin .. nzmax-1
    Wn = W[:N]
    Wn[sel] = (Wn - (zb - dd1[:, None]) * dddt[:, None])[sel]
    hsel = sel[:, :L]
    hn[:N][hsel] = (np.asarray(hnode)[:N] + (zb[:, :L] - zb[:, 1:]) * dd[:, None])[hsel]
    top = uln == 1
    Wn[top, 0] = Wn[top, 0] - np.asarray(water_flux)[:N][top]
    return W, hn


def vert_vel_ale_zlevel(mesh, zbar, hnode, hnode_new, cfl_z, dt, W, hbar, hbar_old, water_flux, min_hnode, lzstar_lev):
    """src/oce_ale.F90:2336-2538 (which_ALE = 'zlevel'), vectorised over the columns: returns (W, hnode_new).  The three
    cases of a cavity-free column -- local zstar over the first lzstar_lev layers (:2367-2454), refill of the subsurface
    layers (:2461-2510), plain zlevel (:2519-2520) -- are masks; the two short recurrences over the layers run as loops
    over k with whole-array operands.  `zbar` is mesh%zbar (nl), `cfl_z` the previous step's CFL_z (nl, Nh)."""
    N, lz = mesh.N, int(lzstar_lev)
    fmax = lambda a, b: np.where(a > b, a, b)            # Fortran MAX / MIN on reals   # noqa: E731
    fmin = lambda a, b: np.where(a < b, a, b)            # noqa: E731
    W, hn = np.array(W, dtype=np.float64), np.array(hnode_new, dtype=np.float64)
    zbar = np.asarray(zbar, dtype=np.float64)
    h = np.asarray(hnode, dtype=np.float64)[:N, :lz]
    cfl = np.asarray(cfl_z, dtype=np.float64)[:N, :lz]
    top = np.asarray(mesh.ulevels_nod2D)[:N] == 1                                  # :2354 (nzmin == 1: layer k sits at index k-1)
    nzmax0 = np.asarray(mesh.nlevels_nod2D_min)[:N] - 1
    dh = np.asarray(hbar, dtype=np.float64)[:N] - np.asarray(hbar_old, dtype=np.float64)[:N]
    rest = (zbar[:lz] - zbar[1:lz + 1])[None, :]
    zero = np.zeros(N)
    # ---- local zstar (:2367-2449)
    A = top & (dh < 0.0) & (h[:, 0] + dh <= rest[0, 0] * min_hnode)
    mx = rest * min_hnode - h
    mx = np.where(mx >= 0.0, 0.0, mx)
    mx = np.where(cfl >= 0.95, 0.0, mx)
    cs = mx.copy()
    cs[:, 1:] = mx[:, 1:] + mx[:, :-1]                                             # the reference's pairwise "cumsum", :2398
    lt = cs < dh[:, None]
    nzA = np.minimum(np.where(lt.any(axis=1), lt.argmax(axis=1) + 1, lz), nzmax0 - 1)      # :2399-2400, :2411
    distrib = np.zeros((N, lz))
    rest_d = dh.copy()
    for k in range(lz):                                                            # :2412-2416
        act = A & (k + 1 <= nzA)
        d = fmax(rest_d, mx[:, k])
        distrib[:, k] = np.where(act, d, 0.0)
        rest_d = np.where(act, fmin(zero, rest_d - d), rest_d)
    integ = np.zeros(N)
    for k in range(lz - 1, -1, -1):                                                # :2438-2449
        act = A & (k + 1 <= nzA)
        integ = np.where(act, integ + distrib[:, k], integ)
        W[:N, k] = np.where(act, W[:N, k] - integ / dt, W[:N, k])
        hn[:N, k] = np.where(act, h[:, k] + distrib[:, k], hn[:N, k])
    # ---- return to zlevel: refill the subsurface layers first (:2461-2510)
    ne = h != rest
    B = top & ~A & (dh > 0.0) & ne[:, 1:].any(axis=1)
    mxB = rest - h
    mxB[:, 0] = 1000.0
    nzB = np.minimum(lz - ne[:, ::-1].argmax(axis=1), nzmax0 - 1)                  # :2482, :2488 (only used where B)
    rest_d = dh.copy()
    integ = np.zeros(N)
    for k in range(lz - 1, -1, -1):
        act = B & (k + 1 <= nzB)
        d = fmin(rest_d, mxB[:, k])
        rest_d = np.where(act, fmax(zero, rest_d - d), rest_d)
        integ = np.where(act, integ + d, integ)
        W[:N, k] = np.where(act, W[:N, k] - integ / dt, W[:N, k])
        hn[:N, k] = np.where(act, h[:, k] + d, hn[:N, k])
    # ---- the normal zlevel case (:2519-2520) and the fresh-water flux (:2527)
    Cn = top & ~A & ~B
    W[:N, 0] = np.where(Cn, W[:N, 0] - dh / dt, W[:N, 0])
    hn[:N, 0] = np.where(Cn, h[:, 0] + dh, hn[:N, 0])
    W[:N, 0] = np.where(top, W[:N, 0] - np.asarray(water_flux, dtype=np.float64)[:N], W[:N, 0])
    return W, hn
