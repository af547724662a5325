// sweep.cuh -- the two stencil sweeps of the explicit step (K5 "pass 1", K6 "pass 2"), version 2.
//
// One thread per particle (cell-sorted order), two phases per batch of stencil columns:
//
//   phase A (FP32 / integer pipes): walk the contiguous particle runs of the stencil columns and
//       test every candidate with a CONSERVATIVE single-precision distance filter on `pf` (position
//       in bucket units, one 16-byte load per candidate); survivors are pushed to a per-thread
//       queue in shared memory (slot-major layout: bank = lane, conflict-free).
//   phase B (FP64 pipe): drain the queue: exact double-precision separation, exact cut-off test,
//       kernel weights and the pair terms.  All lanes of a warp drain together, so the FP64 work
//       runs (nearly) divergence-free instead of being scattered over the candidate loop.
//
// The filter only has to be a superset of the exact predicate: the float coordinates carry an
// error <= 2 ulp_f32(max bucket coordinate) each, which the host turns into a margin on the
// squared cut-off (see filter_radius2 in mphx.cu).  Physics is decided in phase B in fp64.
//
// The reference procedures these kernels replace are listed at the top of kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace mphx {

constexpr int kSweepThreads = 128;
constexpr int kQueueCap = 40; // queue slots per thread (16-byte aligned rows of 128 uint)

struct SweepShared {
    unsigned q[kQueueCap][kSweepThreads];
    int sdx[kMaxStencil], sdy[kMaxStencil], sh[kMaxStencil];
};

// Walks the stencil of particle i in batches of `batch` columns; calls hit(j, dx, dy, dz, r2) in
// phase B for every candidate that passed the fp32 filter (never for j == i).  dx,dy,dz,r2 are the
// exact fp64 separation (minimum image) of j from i.
template <int DIM, class Hit>
__device__ __forceinline__ void sweep(SweepShared &sm, const GridDesc &g, const int *__restrict__ cellStart,
                                      const float4 *__restrict__ pf, const double *__restrict__ X,
                                      const double *__restrict__ Y, const double *__restrict__ Z, int i, bool active,
                                      int key, float filt2, int batch, Hit &&hit)
{
    const int tid = threadIdx.x;
    double xi = 0.0, yi = 0.0, zi = 0.0;
    float fxi = 0.f, fyi = 0.f, fzi = 0.f;
    if (active) {
        xi = X[i]; yi = Y[i]; zi = Z[i];
        const float4 f = pf[i];
        fxi = f.x; fyi = f.y; fzi = f.z;
    }
    int cx, cy, cr, nr;
    if (DIM == 3) { cr = key % g.nz; const int t = key / g.nz; cy = t % g.ny; cx = t / g.ny; nr = g.nz; }
    else { cr = key % g.ny; cx = key / g.ny; cy = 0; nr = g.ny; }
    // does any part of this particle's stencil cross the periodic box?  (rare: then phase B applies
    // the minimum image explicitly; decided per warp so the branch is uniform)
    const int R = g.range;
    bool wraps = (cx - R < 0) || (cx + R >= g.nx) || (cr - R < 0) || (cr + R >= nr);
    if (DIM == 3) wraps = wraps || (cy - R < 0) || (cy + R >= g.ny);
    const bool warp_wraps = __any_sync(0xffffffffu, active && wraps);
    const double W0 = g.W[0], W1 = g.W[1], W2 = g.W[2];
    const double hW0 = 0.5 * W0, hW1 = 0.5 * W1, hW2 = 0.5 * W2;

    int cnt = 0;
    auto drain = [&]() {
        // two queue entries per trip: both position loads are issued before either is used
        for (int s = 0; s < cnt; s += 2) {
            const bool two = s + 1 < cnt;
            const int j0 = (int)sm.q[s][tid];
            const int j1 = two ? (int)sm.q[s + 1][tid] : j0;
            double dx0 = X[j0] - xi, dy0 = Y[j0] - yi, dz0 = Z[j0] - zi;
            double dx1 = X[j1] - xi, dy1 = Y[j1] - yi, dz1 = Z[j1] - zi;
            if (warp_wraps) {
                dx0 = dx0 > hW0 ? dx0 - W0 : (dx0 < -hW0 ? dx0 + W0 : dx0);
                dy0 = dy0 > hW1 ? dy0 - W1 : (dy0 < -hW1 ? dy0 + W1 : dy0);
                dz0 = dz0 > hW2 ? dz0 - W2 : (dz0 < -hW2 ? dz0 + W2 : dz0);
                dx1 = dx1 > hW0 ? dx1 - W0 : (dx1 < -hW0 ? dx1 + W0 : dx1);
                dy1 = dy1 > hW1 ? dy1 - W1 : (dy1 < -hW1 ? dy1 + W1 : dy1);
                dz1 = dz1 > hW2 ? dz1 - W2 : (dz1 < -hW2 ? dz1 + W2 : dz1);
            }
            hit(j0, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0);
            if (two) hit(j1, dx1, dy1, dz1, dx1 * dx1 + dy1 * dy1 + dz1 * dz1);
        }
        cnt = 0;
    };

    // phase A over one contiguous run [jb, je) of candidates, four loads in flight at a time.
    // The caller guarantees room in the queue for the whole run.
    auto scan_run = [&](int jb, int je, float fx, float fy, float fz) {
        for (int j = jb; j < je; j += 4) {
            const float4 far = make_float4(1e18f, 1e18f, 1e18f, 0.f);
            const float4 f0 = __ldg(&pf[j]);
            const float4 f1 = (j + 1 < je) ? __ldg(&pf[j + 1]) : far;
            const float4 f2 = (j + 2 < je) ? __ldg(&pf[j + 2]) : far;
            const float4 f3 = (j + 3 < je) ? __ldg(&pf[j + 3]) : far;
#define MPHX_TEST(f, jj)                                                                           \
    {                                                                                              \
        const float ddx = f.x - fx, ddy = f.y - fy, ddz = f.z - fz;                                \
        const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;                                        \
        if (d2 <= filt2 && (jj) != i) { sm.q[cnt][tid] = (unsigned)(jj); ++cnt; }                  \
    }
            MPHX_TEST(f0, j) MPHX_TEST(f1, j + 1) MPHX_TEST(f2, j + 2) MPHX_TEST(f3, j + 3)
#undef MPHX_TEST
        }
    };
    auto scan_checked = [&](int jb, int je, float fx, float fy, float fz) {
        while (jb < je) { // make room first; a run longer than the queue is scanned in pieces
            const int room = kQueueCap - cnt;
            if (room == 0) { drain(); continue; }
            const int jm = (je - jb <= room) ? je : jb + room;
            scan_run(jb, jm, fx, fy, fz);
            jb = jm;
        }
    };

    const int nsten = g.nsten;
    for (int e0 = 0; e0 < nsten; e0 += batch) {
        const int e1 = (e0 + batch < nsten) ? e0 + batch : nsten;
        if (active && !wraps) {
            // fast path (stencil inside the box): one run per column, bounds of the next column are
            // fetched while the current run is scanned
            auto bounds = [&](int e, int &jb, int &je) {
                const int base = (DIM == 3) ? ((cx + sm.sdx[e]) * g.ny + cy + sm.sdy[e]) * g.nz : (cx + sm.sdx[e]) * g.ny;
                const int h = sm.sh[e];
                jb = cellStart[base + cr - h];
                je = cellStart[base + cr + h + 1];
            };
            int jbn, jen;
            bounds(e0, jbn, jen);
            for (int e = e0; e < e1; ++e) {
                const int jb = jbn, je = jen;
                if (e + 1 < e1) bounds(e + 1, jbn, jen);
                scan_checked(jb, je, fxi, fyi, fzi);
            }
        } else if (active) {
            for (int e = e0; e < e1; ++e) {
                int ccx = cx + sm.sdx[e];
                float fx = fxi, fy = fyi, fz = fzi;
                if (ccx < 0) { ccx += g.nx; fx += (float)g.nx; }
                else if (ccx >= g.nx) { ccx -= g.nx; fx -= (float)g.nx; }
                int base;
                if (DIM == 3) {
                    int ccy = cy + sm.sdy[e];
                    if (ccy < 0) { ccy += g.ny; fy += (float)g.ny; }
                    else if (ccy >= g.ny) { ccy -= g.ny; fy -= (float)g.ny; }
                    base = (ccx * g.ny + ccy) * g.nz;
                } else {
                    base = ccx * g.ny;
                }
                const int h = sm.sh[e];
                const int lo = cr - h, hi = cr + h;
                // in-range part, then the wrapped images
#pragma unroll 1
                for (int seg = 0; seg < 3; ++seg) {
                    int a, b;
                    float shf = 0.f;
                    if (seg == 0) { a = lo < 0 ? 0 : lo; b = hi >= nr ? nr - 1 : hi; }
                    else if (seg == 1) { if (lo >= 0) continue; a = lo + nr; b = nr - 1; shf = (float)nr; }
                    else { if (hi < nr) break; a = 0; b = hi - nr; shf = -(float)nr; }
                    const float fyy = (DIM == 2) ? fy + shf : fy;
                    const float fzz = (DIM == 3) ? fz + shf : fz;
                    scan_checked(cellStart[base + a], cellStart[base + b + 1], fx, fyy, fzz);
                }
            }
        }
        drain();
    }
}

// K5 "pass 1": VolStrainP, DivergenceP -> PressureP (+ DensityA, GravityCenter, PressureA when any
// surface tension is set).  All particle classes (:2320, :2349).
template <int DIM, bool ST>
__global__ void __launch_bounds__(kSweepThreads)
k_pass1_v2(int n, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, float filt2, int batch,
           double *__restrict__ P, double *__restrict__ volStrain, double *__restrict__ divP,
           double *__restrict__ densA, double *__restrict__ gcx, double *__restrict__ gcy, double *__restrict__ gcz,
           double *__restrict__ PA)
{
    __shared__ SweepShared sm;
    for (int e = threadIdx.x; e < g.nsten; e += blockDim.x) { sm.sdx[e] = g.sdx[e]; sm.sdy[e] = g.sdy[e]; sm.sh[e] = g.sh[e]; }
    __syncthreads();
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 < n ? i0 : n - 1;
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const int tflag = p.type[i], ti = real_type(tflag), keyi = p.key[i];
    const bool solid_i = is_structure_type(ti);
    // slab mode: ghosts and parked solids are only ever neighbours; a solid is evaluated by the slab
    // that owns its current column
    bool active = i0 < n && keyi < g.ncells && !(tflag & kGhost);
    if (g.slab && solid_i && active) active = column_owned(g, key_column(g, keyi));
    const double rp2 = ph.rp2, irp = ph.irp, ra2 = ph.ra2, ira = ph.ira;
    double nP = 0.0, dv = 0.0, nA = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0; // nP, dv without their constant factors
    const double *__restrict__ VX = p.vx, *__restrict__ VY = p.vy, *__restrict__ VZ = p.vz;
    const float4 *__restrict__ PF = p.pf;
    sweep<DIM>(sm, g, cellStart, p.pf, p.x, p.y, p.z, i, active, active ? keyi : 0, filt2, batch,
        [&](int j, double dx, double dy, double dz, double r2) {
            if (r2 <= rp2) { // :2333, :2362
                const double rinv = rsqrt(r2);
                const double r = r2 * rinv;
                const double q = 1.0 - r * irp;
                nP += q * q;
                const double ux = VX[j] - vxi, uy = VY[j] - vyi, uz = VZ[j] - vzi;
                dv -= (ux * dx + uy * dy + uz * dz) * rinv * q;
            }
            if (ST && !solid_i && r2 <= ra2) { // :2162, :2195
                const double r = sqrt(r2);
                const double qa = r * ira;
                const double ratio = ph.ratio[ti][__float_as_int(PF[j].w)];
                nA += ratio * (ph.cwa * qa * (1.0 - qa) * (1.0 - qa));
                const double wgv = ratio * (ph.cwg * ((1.0 - qa) * (1.0 - qa))) / ph.r2g * ph.rg;
                g0 += dx * wgv; g1 += dy * wgv; g2 += dz * wgv;
            }
        });
    if (!active) return;
    nP *= ph.cwp;
    dv *= ph.cdp;
    const double vs = nP - ph.n0p;                        // :2339
    const double kappa = (vs < 0.0) ? 0.0 : ph.bulk[ti]; // :2112-2113
    double pr = -ph.lambda[ti] * dv;                      // :2388
    if (vs > 0.0) pr += kappa * vs;                       // :2389-2391
    P[i] = pr; volStrain[i] = vs; divP[i] = dv;
    if (ST) {
        const double da = solid_i ? 0.0 : nA;
        densA[i] = da;
        gcx[i] = solid_i ? 0.0 : g0; gcy[i] = solid_i ? 0.0 : g1; gcz[i] = solid_i ? 0.0 : g2;
        double pa = ph.cofa[ti] * (da - ph.n0a) / ph.l0; // :2219
        if (ph.n0a <= da) pa = 0.0;
        PA[i] = pa;
    }
}

// K6 "pass 2": force sums + gravity + explicit integration (see k_pass2 in kernels.cuh for the
// line-by-line citations; the arithmetic per pair is the same, only the traversal differs).
template <int DIM, bool ST>
__global__ void __launch_bounds__(kSweepThreads)
k_pass2_v2(int n, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, float filt2, int batch,
           const double *__restrict__ P, const double *__restrict__ PA, const double *__restrict__ gcx,
           const double *__restrict__ gcy, const double *__restrict__ gcz, double *__restrict__ ox,
           double *__restrict__ oy, double *__restrict__ oz, double *__restrict__ ovx, double *__restrict__ ovy,
           double *__restrict__ ovz, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
           double *__restrict__ ax, double *__restrict__ ay, double *__restrict__ az, Solid sol,
           double *__restrict__ solbuf)
{
    __shared__ SweepShared sm;
    __shared__ double s_visc[kTypeCount][kTypeCount];
    for (int e = threadIdx.x; e < g.nsten; e += blockDim.x) { sm.sdx[e] = g.sdx[e]; sm.sdy[e] = g.sdy[e]; sm.sh[e] = g.sh[e]; }
    // pair viscosity table with the constant factors of the viscous term folded in:
    // c_d mu_ij V * (-cdv)   (:2505-2512, dwij = -dwvdr)
    for (int e = threadIdx.x; e < kTypeCount * kTypeCount; e += blockDim.x)
        s_visc[e / kTypeCount][e % kTypeCount] = -ph.viscpair[e / kTypeCount][e % kTypeCount] * ph.cdv;
    __syncthreads();
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 < n ? i0 : n - 1;
    const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const int tflag = p.type[i], ti = real_type(tflag), keyi = p.key[i];
    const bool solid_i = is_structure_type(ti);
    bool active = i0 < n && keyi < g.ncells && !(tflag & kGhost);
    if (g.slab && solid_i && active) active = column_owned(g, key_column(g, keyi));
    const double Pi = P[i];
    const double rp2 = ph.rp2, irp = ph.irp, rv2 = ph.rv2, irv = ph.irv;
    const double cpv = ph.cdp * ph.vol; // dwp/dr prefactor times particle volume
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    const double *__restrict__ VX = p.vx, *__restrict__ VY = p.vy, *__restrict__ VZ = p.vz;
    const float4 *__restrict__ PF = p.pf;
    const double *visc_row = s_visc[ti];
    double PAi = 0.0, gi0 = 0.0, gi1 = 0.0, gi2 = 0.0, ai = 0.0;
    if (ST) { PAi = PA[i]; gi0 = gcx[i]; gi1 = gcy[i]; gi2 = gcz[i]; ai = ph.cofa[ti] * ph.cofk * ph.cofk; }
    const double gscale = ph.vol / ph.l0;
    sweep<DIM>(sm, g, cellStart, p.pf, p.x, p.y, p.z, i, active, active ? keyi : 0, filt2, batch,
        [&](int j, double dx, double dy, double dz, double r2) {
            if (solid_i) {
                if (r2 < rp2) { // :2455
                    const int tj = __float_as_int(PF[j].w);
                    if (!is_structure_type(tj)) { // :2447
                        const double rinv = rsqrt(r2);
                        const double c = (Pi + P[j]) * (1.0 - r2 * rinv * irp) * rinv * cpv;
                        F0 += c * dx; F1 += c * dy; F2 += c * dz;
                    }
                }
                return;
            }
            const bool inP = r2 < rp2, inV = r2 < rv2; // :2410, :2496 (strict)
            if (inP || inV) {
                const double rinv = rsqrt(r2);
                const double r = r2 * rinv;
                double c = 0.0;
                if (inP) c = (Pi + P[j]) * (1.0 - r * irp) * rinv * cpv;
                if (inV) {
                    const double ux = VX[j] - vxi, uy = VY[j] - vyi, uz = VZ[j] - vzi;
                    const double ue = (ux * dx + uy * dy + uz * dz) * rinv;
                    c += visc_row[__float_as_int(PF[j].w)] * ue * (1.0 - r * irv) * (rinv * rinv);
                }
                F0 += c * dx; F1 += c * dy; F2 += c * dz;
            }
            if (ST && r2 < ph.ra2) { // :2243, :2285 (RadiusG == RadiusA)
                const int tj = __float_as_int(PF[j].w);
                const double r = sqrt(r2);
                const double rinv = 1.0 / r;
                const double qa = r * ph.ira;
                const double rij = ph.ratio[ti][tj], rji = ph.ratio[tj][ti];
                const double dwa = ph.cwa * (1.0 - qa) * (1.0 - 3.0 * qa) * ph.ira; // dwadr :308
                const double ca = (PAi * (rij * dwa) + PA[j] * (rji * dwa)) * rinv * ph.vol;
                double A0 = ca * dx, A1 = ca * dy, A2 = ca * dz;
                const double wgv = ph.cwg * ((1.0 - qa) * (1.0 - qa));
                const double wij = rij * wgv, wji = rji * wgv;
                const double aj = ai; // Q6: CofA[Property[iP]] for both (:2270, :2275)
                const double gj0 = gcx[j], gj1 = gcy[j], gj2 = gcz[j];
                const double s = gscale * ph.rg / ph.r2g;
                A0 -= (aj * gj0 * wji - ai * gi0 * wij) * s;
                A1 -= (aj * gj1 * wji - ai * gi1 * wij) * s;
                A2 -= (aj * gj2 * wji - ai * gi2 * wij) * s;
                const double dwg = ph.cdg * (1.0 - qa);
                const double dwij = rij * dwg, dwji = rji * dwg;
                const double gr = (aj * gj0 * dwji - ai * gi0 * dwij) * dx + (aj * gj1 * dwji - ai * gi1 * dwij) * dy +
                                  (aj * gj2 * dwji - ai * gi2 * dwij) * dz;
                const double cg = gr * rinv * s;
                A0 -= cg * dx; A1 -= cg * dy; A2 -= cg * dz;
                F0 += A0; F1 += A1; F2 += A2;
            }
        });
    if (i0 >= n) return;
    if (!active) { // ghost / parked / not-owned solid: carried through unchanged (dropped or refreshed next step)
        ox[i] = xi; oy[i] = yi; oz[i] = zi; ovx[i] = vxi; ovy[i] = vyi; ovz[i] = vzi;
        fx[i] = 0.0; fy[i] = 0.0; fz[i] = 0.0; ax[i] = 0.0; ay[i] = 0.0; az[i] = 0.0;
        return;
    }
    // gravity + explicit integration in the reference's operand order (explicitly rounded)
    const double m = ph.mass[ti];
    double nx = xi, ny = yi, nz = zi, nvx = vxi, nvy = vyi, nvz = vzi;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (!is_wall_type(ti)) { // gravity on fluid and solid (:2922-2935)
        F0 = __dadd_rn(F0, __dmul_rn(m, ph.g[0])); F1 = __dadd_rn(F1, __dmul_rn(m, ph.g[1])); F2 = __dadd_rn(F2, __dmul_rn(m, ph.g[2]));
        nvx = __dadd_rn(vxi, __dmul_rn(__ddiv_rn(F0, m), ph.dt)); // :2944-2954
        nvy = __dadd_rn(vyi, __dmul_rn(__ddiv_rn(F1, m), ph.dt));
        nvz = __dadd_rn(vzi, __dmul_rn(__ddiv_rn(F2, m), ph.dt));
        if (!solid_i) { // :1897-1906
            a0 = __ddiv_rn(F0, m); a1 = __ddiv_rn(F1, m); a2 = __ddiv_rn(F2, m);
            nx = __dadd_rn(xi, __dmul_rn(nvx, ph.dt)); ny = __dadd_rn(yi, __dmul_rn(nvy, ph.dt)); nz = __dadd_rn(zi, __dmul_rn(nvz, ph.dt));
        } else {
            const int s = p.id[i] - sol.sb;
            if (solbuf) { // slab mode: published through an all-reduce (all other slabs add zeros)
                const size_t ns = sol.ns;
                solbuf[s] = nvx; solbuf[ns + s] = nvy; solbuf[2 * ns + s] = nvz;
                solbuf[3 * ns + s] = F0; solbuf[4 * ns + s] = F1; solbuf[5 * ns + s] = F2;
            } else {
                sol.vx[s] = nvx; sol.vy[s] = nvy; sol.vz[s] = nvz;
                sol.fx[s] = F0; sol.fy[s] = F1; sol.fz[s] = F2;
            }
        }
    }
    ox[i] = nx; oy[i] = ny; oz[i] = nz; ovx[i] = nvx; ovy[i] = nvy; ovz[i] = nvz;
    fx[i] = F0; fy[i] = F1; fz[i] = F2; ax[i] = a0; ay[i] = a1; az[i] = a2;
}

} // namespace mphx
