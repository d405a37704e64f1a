"""ChainRuleHelper: chain-rule pieces for analytic derivatives of direct-product trial wave functions written in internal
coordinates (bond lengths r, bond angles theta) -- the helper the reference offers to authors of `deriv_function`s
(simulation_utilities/imp_samp_helper.py:10-209; used by call_trl_h2o.py:101-149).

Same class, method names, argument meaning and array shapes as the reference.  `module` is the array module the coordinates
live in: NumPy (host), or `torch` -- then every method runs on the GPU on CUDA tensors (the reference's `xp` idea; CuPy works
through the same code).  Only operations both modules share are used, so a derivative function written on top of this class
can serve the host path (ImpSampManager) and a device-tensor pipeline alike.  The fused water trial wave function on the
device (csrc/pvd_impsamp.cuh, TrialH2OAn) uses the same formulas written out for r1, r2, theta.

Shapes: coords (N, A, 3); dr_dx-like arrays (N, A, 3); stacked over M internal coordinates: (M, N, A, 3) and (M, N)."""

__all__ = ['ChainRuleHelper']


class ChainRuleHelper:
    def __init__(self, coords, module):
        self.cds = coords
        self.xp = module

    # ------------------------------------------------------------------ small array-module neutral helpers
    def _zeros(self):
        return self.cds * 0

    @staticmethod
    def _norm(v):
        return (v * v).sum(axis=-1) ** 0.5

    def _bond(self, a, b):
        return self._norm(self.cds[:, a] - self.cds[:, b])

    def _cos_angle(self, a, b, c):
        """cos of the angle a-b-c (b is the vertex)."""
        u, w = self.cds[:, a] - self.cds[:, b], self.cds[:, c] - self.cds[:, b]
        return (u * w).sum(axis=-1) / (self._norm(u) * self._norm(w))

    # ------------------------------------------------------------------ psi derivatives from internal-coordinate pieces
    def dpsidx(self, dpsi_dr, dr_dx):
        """(psi'_m / psi)(r_m) for m = 0..M-1, shape (M, N), and dr_m/dx, shape (M, N, A, 3) -> grad psi / psi, (N, A, 3):
        sum_m (dr_m/dx) psi'_m/psi (imp_samp_helper.py:10-17)."""
        out = dr_dx[0] * dpsi_dr[0][:, None, None]
        for m in range(1, len(dr_dx)):
            out = out + dr_dx[m] * dpsi_dr[m][:, None, None]
        return out

    def d2psidx2(self, d2psi_dr2, d2r_dx2, dpsi_dr, dr_dx):
        """Second derivatives of a direct product psi = prod_m psi_m(r_m) divided by psi (imp_samp_helper.py:19-39):
        sum_m psi''_m (dr_m/dx)^2 + sum_m psi'_m d2r_m/dx2 + 2 sum_m psi'_m psi'_{m+1} (dr_m/dx)(dr_{m+1}/dx), where the last
        sum runs over cyclic neighbours m, m+1 mod M exactly as the reference's roll does (all pairs for M = 3)."""
        M = len(dr_dx)
        out = self._zeros()
        for m in range(M):
            nxt = (m + 1) % M
            out = out + dr_dx[m] ** 2 * d2psi_dr2[m][:, None, None] + d2r_dx2[m] * dpsi_dr[m][:, None, None] \
                + 2 * (dr_dx[m] * dr_dx[nxt]) * (dpsi_dr[m] * dpsi_dr[nxt])[:, None, None]
        return out

    # ------------------------------------------------------------------ bond lengths
    def dr_dx(self, atm_pair):
        """d r_ab / dx for every Cartesian component: +(x_a - x_b)/r on atom a, -(x_a - x_b)/r on atom b (:41-53)."""
        a, b = atm_pair
        d = self.cds[:, a] - self.cds[:, b]
        unit = d / self._norm(d)[:, None]
        out = self._zeros()
        out[:, a] = unit
        out[:, b] = -unit
        return out

    def d2r_dx2(self, atm_pair, dr_dx=None):
        """d2 r_ab / dx2 = (1 - (dr/dx)^2) / r on both atoms (:55-68)."""
        a, b = atm_pair
        if dr_dx is None:
            dr_dx = self.dr_dx(atm_pair)
        inv_r = (1 / self._bond(a, b))[:, None]
        out = self._zeros()
        out[:, a] = inv_r - inv_r * dr_dx[:, a] ** 2
        out[:, b] = inv_r - inv_r * dr_dx[:, b] ** 2
        return out

    # ------------------------------------------------------------------ cos(theta) of the angle a-b-c (vertex in the middle)
    def dcth_dx(self, atm_pair, cos_theta=None, dr_da=None, dr_dc=None):
        """d cos(theta_abc) / dx (:70-110).  dr_da: dr_dx([a, b]); dr_dc: dr_dx([b, c])."""
        a, b, c = atm_pair
        if cos_theta is None:
            cos_theta = self._cos_angle(a, b, c)
        if dr_da is None or dr_dc is None:
            dr_da, dr_dc = self.dr_dx([a, b]), self.dr_dx([b, c])
        rab, rcb = self._bond(b, a), self._bond(b, c)
        inv_prod = (1 / (rab * rcb))[:, None]
        ca, cc = (cos_theta / rab)[:, None], (cos_theta / rcb)[:, None]
        xa, xb, xc = self.cds[:, a], self.cds[:, b], self.cds[:, c]
        out = self._zeros()
        out[:, a] = (xc - xb) * inv_prod - ca * dr_da[:, a]
        out[:, c] = (xa - xb) * inv_prod - cc * dr_dc[:, c]
        out[:, b] = (2 * xb - xa - xc) * inv_prod - ca * dr_da[:, b] - cc * dr_dc[:, b]
        return out

    def d2cth_dx2(self, atm_pair, cos_theta=None, dr_da=None, dr_dc=None, d2r_da2=None, d2r_dc2=None):
        """d2 cos(theta_abc) / dx2 (:112-165)."""
        a, b, c = atm_pair
        if cos_theta is None:
            cos_theta = self._cos_angle(a, b, c)
        if dr_da is None or dr_dc is None:
            dr_da, dr_dc = self.dr_dx([a, b]), self.dr_dx([b, c])
        if d2r_da2 is None or d2r_dc2 is None:
            d2r_da2, d2r_dc2 = self.d2r_dx2([a, b]), self.d2r_dx2([b, c])
        rab, rcb = self._bond(b, a), self._bond(b, c)
        col = lambda v: v[:, None]
        xa, xb, xc = self.cds[:, a], self.cds[:, b], self.cds[:, c]
        ct = cos_theta
        out = self._zeros()
        out[:, a] = (-2 * (xc - xb)) / col(rab ** 2 * rcb) * dr_da[:, a] + col(2 * ct / rab ** 2) * dr_da[:, a] ** 2 \
            + col(-ct / rab) * d2r_da2[:, a]
        out[:, c] = (-2 * (xa - xb)) / col(rab * rcb ** 2) * dr_dc[:, c] + col(2 * ct / rcb ** 2) * dr_dc[:, c] ** 2 \
            + col(-ct / rcb) * d2r_dc2[:, c]
        mid = 2 * xb - xa - xc
        out[:, b] = col(2 / (rab * rcb)) + col(-ct / rab) * d2r_da2[:, b] + col(-ct / rcb) * d2r_dc2[:, b] \
            + col(2 * ct / rab ** 2) * dr_da[:, b] ** 2 + col(2 * ct / rcb ** 2) * dr_dc[:, b] ** 2 \
            + (-2 * mid / col(rab ** 2 * rcb)) * dr_da[:, b] + (-2 * mid / col(rab * rcb ** 2)) * dr_dc[:, b] \
            + col(2 * ct / (rab * rcb)) * dr_da[:, b] * dr_dc[:, b]
        return out

    # ------------------------------------------------------------------ theta itself
    def dth_dx(self, atm_pair, cos_theta=None, dcth_dx=None, dr_da=None, dr_dc=None):
        """d theta / dx = -1/sqrt(1 - cos^2) * dcos/dx (:167-184)."""
        if dcth_dx is None:
            dcth_dx = self.dcth_dx(atm_pair, dr_da=dr_da, dr_dc=dr_dc)
        if cos_theta is None:
            cos_theta = self._cos_angle(*atm_pair)
        return (-1 / (1 - cos_theta ** 2) ** 0.5)[:, None, None] * dcth_dx

    def d2th_dx2(self, atm_pair, cos_theta=None, dcth_dx=None, dr_da=None, dr_dc=None, d2r_da2=None, d2r_dc2=None):
        """d2 theta / dx2 = (dcos/dx)^2 * (-cos / (1 - cos^2)^1.5) + d2cos/dx2 * (-1 / sqrt(1 - cos^2)) (:186-209)."""
        d2c = self.d2cth_dx2(atm_pair, dr_da=dr_da, dr_dc=dr_dc, d2r_da2=d2r_da2, d2r_dc2=d2r_dc2)
        if cos_theta is None:
            cos_theta = self._cos_angle(*atm_pair)
        if dcth_dx is None:
            dcth_dx = self.dcth_dx(atm_pair, dr_da=dr_da, dr_dc=dr_dc)
        s2 = 1 - cos_theta ** 2
        return dcth_dx ** 2 * (-cos_theta / s2 ** 1.5)[:, None, None] + d2c * (-1 / s2 ** 0.5)[:, None, None]
