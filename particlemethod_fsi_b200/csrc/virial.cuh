// virial.cuh -- calculateVirialStressAtParticle (src/main.cpp:3077-3318), SURVEY.md 8(f) N2.
//
// The reference evaluates it on VTK-output steps only (:671-673), over its Neighbor lists -- built by the
// step's calculateNeighbor on the PRE-step positions with the (MaxRadius+MARGIN) cut-off -- but with the
// separations, velocities and pressures of the state AFTER the step.  So this kernel walks a private bucket
// structure over the pre-step positions with the reference's bit-exact list predicate (like
// k_neighbors_exact) and evaluates the four pair terms on the current state:
//   pressure P   :3095-3131   f_ij = P_i  grad wp  V                         (r^2 < RadiusP^2)
//   pressure A   :3140-3183   f_ij = PA_i ratio_ij grad wa V                 (r^2 < RadiusA^2)
//   viscosity    :3186-3232   f_ij = c_d mu_ij (u_ij.e_ij) e_ij (-wv')/r V,  weight 1/2   (r^2 < RadiusV^2)
//   diffuse interface, two terms :3235-3303 (GravityCenter of i)             (r^2 < RadiusG^2)
//   stress_i += coef f_ij (x) x_ij / V;   VirialPressure = -tr/d  (:3309-3316)
// It is an output-step diagnostic, not part of the step: one thread per particle, plain fp64 (the sums run
// in bucket order, not in the reference's list order: 1e-10, not bit-exact).
#pragma once
#include "kernels.cuh"

namespace mphx {

// current state by slot: fluid / wall from the sorted arrays, solids from the solid arrays (they are integrated there)
__global__ void k_current_state(int n, Particles p, Solid sol, double *__restrict__ cx, double *__restrict__ cy, double *__restrict__ cz,
                                double *__restrict__ cvx, double *__restrict__ cvy, double *__restrict__ cvz)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int t = p.type[q];
    if (is_structure_type(t) && !(t & kGhost)) {
        const int s = p.id[q] - sol.sb;
        cx[q] = sol.x[s]; cy[q] = sol.y[s]; cz[q] = sol.z[s]; cvx[q] = sol.vx[s]; cvy[q] = sol.vy[s]; cvz[q] = sol.vz[s];
    } else {
        cx[q] = p.x[q]; cy[q] = p.y[q]; cz[q] = p.z[q]; cvx[q] = p.vx[q]; cvy[q] = p.vy[q]; cvz[q] = p.vz[q];
    }
}
__global__ void k_iota(int n, int *__restrict__ a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) a[q] = q;
}

struct VirialIn {
    const double *sx, *sy, *sz; // pre-step positions, bucket-sorted (the private structure)
    const int *sslot, *skey;    // slot of each sorted entry, its bucket
    const int *cellStart;
    const double *cx, *cy, *cz, *cvx, *cvy, *cvz; // current state, by slot
    const int *type, *id;       // by slot
    const double *P, *PA, *gcx, *gcy, *gcz; // by slot (PA, gc: nullptr without surface tension)
    double *out9, *outp;        // [N][3][3], [N] in ORIGINAL particle order
};

template <int DIM>
__global__ void __launch_bounds__(128)
k_virial(int n, VirialIn in, GridDesc g, Phys ph, double cutoff2, int surface_tension)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int i = in.sslot[q];
    const int tflag = in.type[i];
    if (in.skey[q] >= g.ncells || (tflag & kGhost)) return;
    const int ti = real_type(tflag);
    const double xi = in.sx[q], yi = in.sy[q], zi = in.sz[q];
    const double cxi = in.cx[i], cyi = in.cy[i], czi = in.cz[i];
    const double vxi = in.cvx[i], vyi = in.cvy[i], vzi = in.cvz[i];
    const double Pi = in.P[i];
    const double PAi = surface_tension ? in.PA[i] : 0.0;
    const double gi[3] = {surface_tension ? in.gcx[i] : 0.0, surface_tension ? in.gcy[i] : 0.0, surface_tension ? in.gcz[i] : 0.0};
    const double ai = ph.cofa[ti] * ph.cofk * ph.cofk;
    const double gscale = ph.rg / ph.r2g * (ph.vol / ph.l0);
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for_each_candidate<DIM>(g, in.cellStart, in.sx, in.sy, in.sz, in.skey[q], xi, yi, zi,
        [&](int jq, double, double, double, double) {
            if (jq == q) return;
            // the reference's list predicate on the pre-step positions (:1759-1772)
            const double p0 = minimg_exact(in.sx[jq], xi, g.W[0]), p1 = minimg_exact(in.sy[jq], yi, g.W[1]), p2 = minimg_exact(in.sz[jq], zi, g.W[2]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, p0), __dmul_rn(p1, p1)), __dmul_rn(p2, p2));
            if (!(d2 <= cutoff2)) return;
            const int j = in.sslot[jq];
            const int tj = real_type(in.type[j]);
            const double x[3] = {minimg_exact(in.cx[j], cxi, g.W[0]), minimg_exact(in.cy[j], cyi, g.W[1]), minimg_exact(in.cz[j], czi, g.W[2])};
            const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
            double f[3] = {0.0, 0.0, 0.0}; // sum of coef * f_ij over the terms (the common factor x_ij / V follows)
            const double r = sqrt(r2), rinv = 1.0 / r;
            if (ph.rp2 - r2 > 0) { // :3108
                const double c = Pi * (ph.cdp * (1.0 - r * ph.irp)) * rinv * ph.vol;
                for (int d = 0; d < 3; ++d) f[d] += c * x[d];
            }
            if (ph.rv2 - r2 > 0) { // :3199, weight 0.5 :3226
                const double ue = ((in.cvx[j] - vxi) * x[0] + (in.cvy[j] - vyi) * x[1] + (in.cvz[j] - vzi) * x[2]) * rinv;
                const double c = 0.5 * ph.viscpair[ti][tj] * ue * (-(ph.cdv * (1.0 - r * ph.irv))) * rinv * rinv;
                for (int d = 0; d < 3; ++d) f[d] += c * x[d];
            }
            if (surface_tension && ph.ra2 - r2 > 0) { // :3153, :3248, :3270 (RadiusG == RadiusA)
                const double ratio = ph.ratio[ti][tj];
                const double qa = r * ph.ira;
                const double ca = PAi * (ratio * (ph.cwa * (1.0 - qa) * (1.0 - 3.0 * qa) * ph.ira)) * rinv * ph.vol;
                const double wgv = ratio * (ph.cwg * ((1.0 - qa) * (1.0 - qa)));
                const double dwg = ratio * (ph.cdg * (1.0 - qa));
                const double gr = -(gi[0] * x[0] + gi[1] * x[1] + gi[2] * x[2]);
                const double cg = -ai * gr * dwg * rinv * gscale;
                for (int d = 0; d < 3; ++d) f[d] += ca * x[d] + ai * gi[d] * wgv * gscale + cg * x[d];
            }
            const double iv = 1.0 / ph.vol;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) S[a][b] += f[a] * x[b] * iv;
        });
    const size_t o = (size_t)in.id[i];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) in.out9[9 * o + 3 * a + b] = S[a][b];
    in.outp[o] = (DIM == 2) ? -1.0 / 2.0 * (S[0][0] + S[1][1]) : -1.0 / 3.0 * (S[0][0] + S[1][1] + S[2][2]); // :3309-3316
}

} // namespace mphx
