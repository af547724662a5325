// kernels.cuh -- hand-written sm_100a kernels of the explicit MPH / total-Lagrangian step.
//
// Layout: every per-particle field is a cell-sorted structure of arrays (one double/int array per
// component).  Cells are the reference's buckets (width = one particle spacing, CellId x-major /
// z-minor, src/main.cpp:123-125,1414) so the sort key IS the reference's CellIndex; particles of
// the 2h+1 cells of a stencil column that are adjacent along the minor ("run") axis are contiguous
// in memory, so a particle's (2h+1)^d-cell neighbourhood is (2h+1)^(d-1) contiguous runs.
//
// Reference procedures replaced (src/main.cpp): calculateWall :2963, calculatePeriodicBoundary :3322,
// resetForce :2085, resetAccel :2892, calculateNeighbor :1662, calculateDensityA :2141,
// calculateGravityCenter :2174, calculateDensityP :2314, calculateDivergenceP :2343,
// calculatePhysicalCoefficients :2099, calculatePressureP :2381, calculatePressureA :2212,
// calculateDiffuseInterface :2261, calculateViscosityV :2478, calculateGravity :2917,
// calculateInterfaceForce :2427, calculateAcceleration :2938, calculateConvection :1892,
// calculateInitialNeighbor :1497, calculateNormalizer :2544, calculateElasticDeformationVector :2673,
// calculateStress :2756, calculateStressForce :2812, updateElasticPosition :1910.
#pragma once
#include <cuda_runtime.h>

namespace mphx {

constexpr int kMaxStencil = 128;
constexpr int kTypeCount = 6;

__host__ __device__ inline bool is_structure_type(int t) { return t >= 2 && t < 4; }
__host__ __device__ inline bool is_fluid_type(int t) { return t >= 0 && t < 2; }
__host__ __device__ inline bool is_wall_type(int t) { return t >= 4 && t < 6; }

struct GridDesc {
    int dim;        // 2 or 3
    int nx, ny, nz; // bucket counts (nz = 1 in 2D)
    int ncells;
    int nsten; // stencil columns
    int range; // ceil((MaxRadius+MARGIN)/CellWidth)  src/main.cpp:1744
    int pad;
    double mn[3], W[3], cellw;
    // stencil columns.  3D: (dx,dy) with half-length sh along z.  2D: dx with half-length sh along y.
    signed char sdx[kMaxStencil], sdy[kMaxStencil], sh[kMaxStencil];
};

struct Phys {
    double dt, vol, l0;
    double rp2, irp, cwp, cdp; // wp = cwp*(1-r/RP)^2, dwp/dr = cdp*(1-r/RP)
    double rv2, irv, cdv;      // dwv/dr = cdv*(1-r/RV)
    double ra2, ira, cwa, cwg, cdg, r2g, rg; // surface-tension kernels (RadiusG = RadiusA)
    double n0p, n0a, cofk;
    double mass[kTypeCount], inv_density[kTypeCount], bulk[kTypeCount], lambda[kTypeCount];
    double cofa[kTypeCount];
    double viscpair[kTypeCount][kTypeCount]; // c_d * mu_ij * V,  c_d = 8 (2D) / 10 (3D)
    double ratio[kTypeCount][kTypeCount];
    double g[3];
};

struct WallMotion {
    int active; // Time < 0.2  (src/main.cpp:3037)
    int pad;
    double dt;
    double center[kTypeCount][3], vel[kTypeCount][3], omega[kTypeCount][3], R[kTypeCount][3][3];
};

// cell-sorted particle arrays
struct Particles {
    double *x, *y, *z, *vx, *vy, *vz;
    int *type, *id, *key;
    float4 *pf; // (position - DomainMin)/CellWidth in fp32 + type bits: input of the sweep's fp32 filter
};

// total-Lagrangian solid, static order (solid-local index s = original id - sb)
struct Solid {
    int ns, sb;
    double *x, *y, *z, *vx, *vy, *vz, *x0, *y0, *z0, *fx, *fy, *fz;
    double *Linv, *Fm, *E, *S, *Pk; // 9 planes of ns doubles each: M[k*ns+s], k = 3*row+col
    double *lam, *mu;
    int *type;
    int *off, *nbr;   // InitialStructureNeighbor as CSR (solid-local ids, rows in the reference's list order)
    int *roff, *rnbr; // transpose: rows = particles that list s (ascending)
    // The sub-step kernels read the lists in ELL layout (entry kk of row s at [kk*ns + s]: coalesced
    // across the threads of a warp) together with static per-pair data of the reference
    // configuration computed once: x0_ij and weight(x0_ij).
    int *len, *rlen;                 // row lengths
    int *enbr, *ernbr;               // ELL neighbour ids (own rows / transposed rows)
    double *d0x, *d0y, *d0z, *w;     // own rows
    double *rd0x, *rd0y, *rd0z, *rw; // transposed: x0_js as row j computes it
    double *ux, *uy, *uz;            // displacement u = minimg(x - x0) of the current positions
};

// ------------------------------------------------------------------------------------------------
// bit-exact helpers (IEEE, no FMA contraction): the reference's Mod macro (:98) and CellId (:123-125)
__device__ __forceinline__ double mod_exact(double x, double w)
{
    return __dsub_rn(x, __dmul_rn(w, floor(__ddiv_rn(x, w))));
}
__device__ __forceinline__ double minimg_exact(double xj, double xi, double W)
{
    const double h = __dmul_rn(0.5, W);
    return __dsub_rn(mod_exact(__dadd_rn(__dsub_rn(xj, xi), h), W), h);
}
__device__ __forceinline__ int cell_coord_exact(double x, double mn, double cw, int n)
{
    int c = __double2int_rz(floor(__ddiv_rn(__dsub_rn(x, mn), cw))) % n; // :1671
    return (c % n + n) % n;                                               // CellId wrap
}
__device__ __forceinline__ int cell_key(const GridDesc &g, double x, double y, double z)
{
    const int cx = cell_coord_exact(x, g.mn[0], g.cellw, g.nx);
    const int cy = cell_coord_exact(y, g.mn[1], g.cellw, g.ny);
    if (g.dim == 2) return cx * g.ny + cy;
    const int cz = cell_coord_exact(z, g.mn[2], g.cellw, g.nz);
    return (cx * g.ny + cy) * g.nz + cz;
}
// plain (FMA-allowed) minimum image for the 1e-10 paths
__device__ __forceinline__ double minimg(double d, double W)
{
    const double h = 0.5 * W;
    const double t = d + h;
    return (t - W * floor(t / W)) - h;
}

// ------------------------------------------------------------------------------------------------
// K0/K1: pre-step.  wall kinematics (t<0.2), periodic wrap, bucket key, per-bucket count + slot.
// Solids take their state from the solid arrays (they are integrated there).
__global__ void k_prestep(int n, Particles p, Solid sol, GridDesc g, WallMotion wm, int do_wrap,
                          int *__restrict__ cellCount, int *__restrict__ slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = p.type[i];
    double x = p.x[i], y = p.y[i], z = p.z[i];
    int s = -1;
    if (is_structure_type(t)) {
        s = p.id[i] - sol.sb;
        x = sol.x[s]; y = sol.y[s]; z = sol.z[s];
        p.vx[i] = sol.vx[s]; p.vy[i] = sol.vy[s]; p.vz[i] = sol.vz[s];
    } else if (wm.active && is_wall_type(t)) { // :3036-3060, operand order kept (bit-exact)
        const double r0 = __dsub_rn(x, wm.center[t][0]), r1 = __dsub_rn(y, wm.center[t][1]),
                     r2 = __dsub_rn(z, wm.center[t][2]);
        const double(*R)[3] = wm.R[t];
        const double q0 = __dadd_rn(__dadd_rn(__dmul_rn(R[0][0], r0), __dmul_rn(R[0][1], r1)), __dmul_rn(R[0][2], r2));
        const double q1 = __dadd_rn(__dadd_rn(__dmul_rn(R[1][0], r0), __dmul_rn(R[1][1], r1)), __dmul_rn(R[1][2], r2));
        const double q2 = __dadd_rn(__dadd_rn(__dmul_rn(R[2][0], r0), __dmul_rn(R[2][1], r1)), __dmul_rn(R[2][2], r2));
        const double *w = wm.omega[t], *V = wm.vel[t];
        p.vx[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[1], q2), __dmul_rn(w[2], q1)), V[0]);
        p.vy[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[2], q0), __dmul_rn(w[0], q2)), V[1]);
        p.vz[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[0], q1), __dmul_rn(w[1], q0)), V[2]);
        x = __dadd_rn(__dadd_rn(q0, wm.center[t][0]), __dmul_rn(V[0], wm.dt));
        y = __dadd_rn(__dadd_rn(q1, wm.center[t][1]), __dmul_rn(V[1], wm.dt));
        z = __dadd_rn(__dadd_rn(q2, wm.center[t][2]), __dmul_rn(V[2], wm.dt));
    }
    if (do_wrap) { // :3330
        x = __dadd_rn(mod_exact(__dsub_rn(x, g.mn[0]), g.W[0]), g.mn[0]);
        y = __dadd_rn(mod_exact(__dsub_rn(y, g.mn[1]), g.W[1]), g.mn[1]);
        z = __dadd_rn(mod_exact(__dsub_rn(z, g.mn[2]), g.W[2]), g.mn[2]);
    }
    p.x[i] = x; p.y[i] = y; p.z[i] = z;
    if (s >= 0) {
        sol.x[s] = x; sol.y[s] = y; sol.z[s] = z;
        sol.ux[s] = minimg_exact(x, sol.x0[s], g.W[0]); // :2712, from the (wrapped) position the sub-steps start from
        sol.uy[s] = minimg_exact(y, sol.y0[s], g.W[1]);
        sol.uz[s] = minimg_exact(z, sol.z0[s], g.W[2]);
    }
    const int k = cell_key(g, x, y, z);
    p.key[i] = k;
    slot[i] = atomicAdd(&cellCount[k], 1);
}

// ------------------------------------------------------------------------------------------------
// K2: exclusive scan of the bucket counts (three small kernels; int32)
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4; // per thread
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int ws = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - ws; // exclusive prefix of warp sums
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const int r = warp_sums[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const int *__restrict__ in, int n, int *__restrict__ blockSums)
{
    __shared__ int total;
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in[base + k];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_top(int *__restrict__ blockSums, int nb)
{
    __shared__ int total;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += kScanThreads) {
        const int idx = b0 + threadIdx.x;
        const int v = idx < nb ? blockSums[idx] : 0;
        const int ex = block_exclusive_scan(v, &total);
        const int c = carry;
        if (idx < nb) blockSums[idx] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ blockPrefix,
                                                             int *__restrict__ out /* n+1 */)
{
    __shared__ int total;
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int ex = block_exclusive_scan(s, &total) + blockPrefix[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

// K3: provisional bucket order (arrival order inside a bucket is arbitrary -> fixed up in K4)
__global__ void k_scatter_index(int n, const int *__restrict__ key, const int *__restrict__ slot,
                                const int *__restrict__ cellStart, int *__restrict__ tmpIdx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tmpIdx[cellStart[key[i]] + slot[i]] = i;
}

// K4: permute the SoA into bucket order; inside a bucket particles are ordered by original id, which
// makes the layout (and therefore every floating-point sum) independent of atomic arrival order.
__global__ void k_permute(int n, Particles src, Particles dst, const int *__restrict__ cellStart,
                          const int *__restrict__ tmpIdx, GridDesc g)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int s = tmpIdx[q];
    const int k = src.key[s];
    const int b = cellStart[k], e = cellStart[k + 1];
    if (e - b > 1) {
        const int want = q - b;
        for (int a = b; a < e; ++a) {
            const int sa = tmpIdx[a];
            const int ida = src.id[sa];
            int rank = 0;
            for (int c = b; c < e; ++c) rank += (src.id[tmpIdx[c]] < ida);
            if (rank == want) { s = sa; break; }
        }
    }
    const double x = src.x[s], y = src.y[s], z = src.z[s];
    const int t = src.type[s];
    dst.x[q] = x; dst.y[q] = y; dst.z[q] = z;
    dst.vx[q] = src.vx[s]; dst.vy[q] = src.vy[s]; dst.vz[q] = src.vz[s];
    dst.type[q] = t; dst.id[q] = src.id[s]; dst.key[q] = k;
    const double icw = 1.0 / g.cellw;
    dst.pf[q] = make_float4((float)((x - g.mn[0]) * icw), (float)((y - g.mn[1]) * icw), (float)((z - g.mn[2]) * icw),
                            __int_as_float(t));
}

// ------------------------------------------------------------------------------------------------
// stencil traversal: calls f(j, dx, dy, dz, r2) for every particle j of the stencil buckets of the
// bucket `key` (including i itself).  Periodic images are handled by shifting x_i per run segment.
template <int DIM, class F>
__device__ __forceinline__ void for_each_candidate(const GridDesc &g, const int *__restrict__ cellStart,
                                                   const double *__restrict__ X, const double *__restrict__ Y,
                                                   const double *__restrict__ Z, int key, double xi, double yi,
                                                   double zi, F &&f)
{
    int cx, cy, cr, nr; // cr/nr: coordinate / count along the run axis
    if (DIM == 3) {
        cr = key % g.nz; const int t = key / g.nz; cy = t % g.ny; cx = t / g.ny; nr = g.nz;
    } else {
        cr = key % g.ny; cx = key / g.ny; cy = 0; nr = g.ny;
    }
    for (int e = 0; e < g.nsten; ++e) {
        int ccx = cx + g.sdx[e];
        double xs = xi;
        if (ccx < 0) { ccx += g.nx; xs = xi + g.W[0]; }
        else if (ccx >= g.nx) { ccx -= g.nx; xs = xi - g.W[0]; }
        int base;
        double ys = yi, zs = zi;
        if (DIM == 3) {
            int ccy = cy + g.sdy[e];
            if (ccy < 0) { ccy += g.ny; ys = yi + g.W[1]; }
            else if (ccy >= g.ny) { ccy -= g.ny; ys = yi - g.W[1]; }
            base = (ccx * g.ny + ccy) * g.nz;
        } else {
            base = ccx * g.ny;
        }
        const int h = g.sh[e];
        const int lo = cr - h, hi = cr + h;
        // up to three segments: wrapped-low image, in-range part, wrapped-high image
#pragma unroll 1
        for (int seg = 0; seg < 3; ++seg) {
            int a, b;
            double sh = 0.0;
            if (seg == 0) { if (lo >= 0) continue; a = lo + nr; b = nr - 1; sh = g.W[DIM - 1]; }
            else if (seg == 1) { a = lo < 0 ? 0 : lo; b = hi >= nr ? nr - 1 : hi; }
            else { if (hi < nr) continue; a = 0; b = hi - nr; sh = -g.W[DIM - 1]; }
            const double yy = (DIM == 2) ? ys + sh : ys;
            const double zz = (DIM == 3) ? zs + sh : zs;
            const int jb = cellStart[base + a], je = cellStart[base + b + 1];
            for (int j = jb; j < je; ++j) {
                const double dx = X[j] - xs, dy = Y[j] - yy, dz = Z[j] - zz;
                const double r2 = dx * dx + dy * dy + dz * dz;
                f(j, dx, dy, dz, r2);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5 "pass 1": VolStrainP, DivergenceP -> PressureP (+ DensityA, GravityCenter, PressureA when any
// surface tension is set).  One thread per particle, all particle classes (:2320, :2349).
template <int DIM, bool ST>
__global__ void __launch_bounds__(128)
k_pass1(int n, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, double *__restrict__ P,
        double *__restrict__ volStrain, double *__restrict__ divP, double *__restrict__ densA,
        double *__restrict__ gcx, double *__restrict__ gcy, double *__restrict__ gcz, double *__restrict__ PA)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const int ti = p.type[i];
    const bool solid_i = is_structure_type(ti);
    double nP = 0.0, dv = 0.0, nA = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    const double *__restrict__ VX = p.vx, *__restrict__ VY = p.vy, *__restrict__ VZ = p.vz;
    const int *__restrict__ TY = p.type;
    for_each_candidate<DIM>(g, cellStart, p.x, p.y, p.z, p.key[i], xi, yi, zi,
        [&](int j, double dx, double dy, double dz, double r2) {
            if (j == i) return;
            if (r2 <= ph.rp2) { // :2333, :2362
                const double r = sqrt(r2);
                const double q = 1.0 - r * ph.irp;
                nP += ph.cwp * (q * q);
                const double ux = VX[j] - vxi, uy = VY[j] - vyi, uz = VZ[j] - vzi;
                dv -= (ux * dx + uy * dy + uz * dz) / r * (ph.cdp * q);
            }
            if (ST && !solid_i && r2 <= ph.ra2) { // :2162, :2195
                const double r = sqrt(r2);
                const double qa = r * ph.ira;
                const double ratio = ph.ratio[ti][TY[j]];
                nA += ratio * (ph.cwa * qa * (1.0 - qa) * (1.0 - qa));
                const double wgv = ratio * (ph.cwg * ((1.0 - qa) * (1.0 - qa))) / ph.r2g * ph.rg;
                g0 += dx * wgv; g1 += dy * wgv; g2 += dz * wgv;
            }
        });
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

// ------------------------------------------------------------------------------------------------
// K6 "pass 2": force sums + gravity + explicit integration.
//   fluid/wall i : pressure (:2397-2424), attractive pressure (:2228-2258), diffuse interface
//                  (:2268-2311), viscosity (:2483-2521)
//   solid i      : fluid->solid interface force (:2442-2472)
//   then gravity (:2922-2935), v += F/m dt (:2943-2955), fluid a += F/m, x += v dt (:1897-1906).
// New x/v go to the `out` arrays (the inputs are still being read by neighbouring threads).
template <int DIM, bool ST>
__global__ void __launch_bounds__(128)
k_pass2(int n, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, const double *__restrict__ P,
        const double *__restrict__ PA, const double *__restrict__ gcx, const double *__restrict__ gcy,
        const double *__restrict__ gcz, double *__restrict__ ox, double *__restrict__ oy, double *__restrict__ oz,
        double *__restrict__ ovx, double *__restrict__ ovy, double *__restrict__ ovz, double *__restrict__ fx,
        double *__restrict__ fy, double *__restrict__ fz, double *__restrict__ ax, double *__restrict__ ay,
        double *__restrict__ az, Solid sol)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const int ti = p.type[i];
    const bool solid_i = is_structure_type(ti);
    const double Pi = P[i];
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    const double *__restrict__ VX = p.vx, *__restrict__ VY = p.vy, *__restrict__ VZ = p.vz;
    const int *__restrict__ TY = p.type;
    double PAi = 0.0, gi0 = 0.0, gi1 = 0.0, gi2 = 0.0, ai = 0.0;
    if (ST) { PAi = PA[i]; gi0 = gcx[i]; gi1 = gcy[i]; gi2 = gcz[i]; ai = ph.cofa[ti] * ph.cofk * ph.cofk; }
    const double gscale = ph.vol / ph.l0;
    for_each_candidate<DIM>(g, cellStart, p.x, p.y, p.z, p.key[i], xi, yi, zi,
        [&](int j, double dx, double dy, double dz, double r2) {
            if (j == i) return;
            if (solid_i) {
                if (r2 < ph.rp2) { // :2455
                    const int tj = TY[j];
                    if (is_structure_type(tj)) return; // :2447
                    const double r = sqrt(r2);
                    const double c = (Pi + P[j]) * (ph.cdp * (1.0 - r * ph.irp)) / r * ph.vol;
                    F0 += c * dx; F1 += c * dy; F2 += c * dz;
                }
                return;
            }
            const bool inP = r2 < ph.rp2, inV = r2 < ph.rv2; // :2410, :2496 (strict)
            if (inP || inV) {
                const double r = sqrt(r2);
                const double rinv = 1.0 / r;
                double c = 0.0;
                if (inP) c = (Pi + P[j]) * (ph.cdp * (1.0 - r * ph.irp)) * rinv * ph.vol;
                if (inV) {
                    const double ux = VX[j] - vxi, uy = VY[j] - vyi, uz = VZ[j] - vzi;
                    const double ue = (ux * dx + uy * dy + uz * dz) * rinv;
                    const double dwij = -(ph.cdv * (1.0 - r * ph.irv));
                    c += ph.viscpair[ti][TY[j]] * ue * dwij * rinv * rinv;
                }
                F0 += c * dx; F1 += c * dy; F2 += c * dz;
            }
            if (ST && r2 < ph.ra2) { // :2243, :2285 (RadiusG == RadiusA)
                const int tj = TY[j];
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
    // gravity + explicit integration in the reference's operand order (explicitly rounded, so a
    // particle whose force sum is exact -- e.g. a solid far from any fluid -- moves bit-identically)
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
            sol.vx[s] = nvx; sol.vy[s] = nvy; sol.vz[s] = nvz;
            sol.fx[s] = F0; sol.fy[s] = F1; sol.fz[s] = F2;
        }
    }
    ox[i] = nx; oy[i] = ny; oz[i] = nz; ovx[i] = nvx; ovy[i] = nvy; ovz[i] = nvz;
    fx[i] = F0; fy[i] = F1; fz[i] = F2; ax[i] = a0; ay[i] = a1; az[i] = a2;
}

// ------------------------------------------------------------------------------------------------
// neighbour SETS with the reference's bit-exact predicate (calculateNeighbor :1759-1772 and
// calculateInitialNeighbor :1601-1616).  MODE 0 = count, 1 = fill (+ sort row ascending).
//   structure_only: rows and entries restricted to structure particles (initial lists)
//   xy_only: 2D initial lists use two components (:1602-1605)
// Rows are indexed by (id - row_base); entries are (id_j - row_base).
template <int DIM, int MODE>
__global__ void __launch_bounds__(128)
k_neighbors_exact(int n, const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                  const int *__restrict__ type, const int *__restrict__ id, const int *__restrict__ key,
                  const int *__restrict__ cellStart, GridDesc g, double cutoff2, int structure_only, int xy_only,
                  int row_base, int *__restrict__ counts, const long long *__restrict__ offsets,
                  int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (structure_only && !is_structure_type(type[i])) return;
    const double xi = X[i], yi = Y[i], zi = Z[i];
    const int row = id[i] - row_base;
    int cnt = 0;
    int *dst = (MODE == 1) ? out + offsets[row] : nullptr;
    for_each_candidate<DIM>(g, cellStart, X, Y, Z, key[i], xi, yi, zi,
        [&](int j, double, double, double, double) {
            if (j == i) return;
            if (structure_only && !is_structure_type(type[j])) return;
            const double q0 = minimg_exact(X[j], xi, g.W[0]);
            const double q1 = minimg_exact(Y[j], yi, g.W[1]);
            const double q2 = xy_only ? 0.0 : minimg_exact(Z[j], zi, g.W[2]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)), __dmul_rn(q2, q2));
            if (d2 <= cutoff2) {
                if (MODE == 1) dst[cnt] = id[j] - row_base;
                ++cnt;
            }
        });
    if (MODE == 0) counts[row] = cnt;
    else {
        for (int a = 1; a < cnt; ++a) { // insertion sort, rows are short (<= ~80)
            const int v = dst[a];
            int b = a - 1;
            while (b >= 0 && dst[b] > v) { dst[b + 1] = dst[b]; --b; }
            dst[b + 1] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// total-Lagrangian solid.  M(k) = plane k of a 3x3 SoA tensor.
//
// The reference-configuration lists are static and hold at most one particle per bucket, so the
// reference's accumulation ORDER can be reproduced exactly (rows are stored in its stencil order).
// The solid kernels therefore use explicitly rounded operations in the reference's operand order
// (no FMA contraction): given identical inputs they return the reference's bits.  This matters
// because E = (F^T F - I)/2 cancels ~5 digits at small strain, so any re-association shows up at
// 1e-11 relative per sub-step.
#define MPHX_T(M, r, c, s, ns) (M)[(size_t)(3 * (r) + (c)) * (ns) + (s)]

namespace ex {
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
} // namespace ex

// weight() :268-295 -- always the pressure normaliser Swp; 2D ignores the third component.
// cw = (1.0/Swp)*(1.0/radius^d) evaluated on the host in that order.
template <int DIMS>
__device__ __forceinline__ double tl_weight(double x0, double x1, double x2, double radius, double cw)
{
    double r2 = ex::add(ex::mul(x0, x0), ex::mul(x1, x1)); // 0.0 + x0*x0 is exact
    if (DIMS == 3) r2 = ex::add(r2, ex::mul(x2, x2));
    const double q = ex::div(sqrt(r2), radius);
    const double omq = ex::sub(1.0, q);
    return ex::mul(cw, ex::mul(omq, omq));
}

// calculateNormalizer :2544-2653 (once)
template <int DIMS>
__global__ void k_solid_normalizer(Solid so, double W0, double W1, double W2, double radius, double cw)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const double xi0 = so.x0[s], yi0 = so.y0[s], zi0 = so.z0[s];
    for (int k = so.off[s]; k < so.off[s + 1]; ++k) {
        const int j = so.nbr[k];
        // Q2: three components are accumulated even in 2D (the reference tests a misspelt macro)
        const double d[3] = {minimg_exact(so.x0[j], xi0, W0), minimg_exact(so.y0[j], yi0, W1), minimg_exact(so.z0[j], zi0, W2)};
        const double w = tl_weight<DIMS>(d[0], d[1], d[2], radius, cw);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) A[a][b] = ex::add(A[a][b], ex::mul(ex::mul(w, d[a]), d[b]));
    }
    using namespace ex;
    if (DIMS == 2) { // :2592-2621
        const double a = A[0][0], b = A[0][1], c = A[1][0], d = A[1][1];
        const double det = sub(mul(a, d), mul(b, c));
        if (det != 0.0) { A[0][0] = div(d, det); A[0][1] = div(-b, det); A[1][0] = div(-c, det); A[1][1] = div(a, det); }
        else { A[0][0] = 1.0; A[0][1] = 0.0; A[1][0] = 0.0; A[1][1] = 1.0; }
    } else { // :2624-2650
        const double det = add(sub(mul(A[0][0], sub(mul(A[1][1], A[2][2]), mul(A[1][2], A[2][1]))),
                                   mul(A[0][1], sub(mul(A[1][0], A[2][2]), mul(A[1][2], A[2][0])))),
                               mul(A[0][2], sub(mul(A[1][0], A[2][1]), mul(A[1][1], A[2][0]))));
        if (det != 0.0) {
            double B[3][3];
            B[0][0] = sub(mul(A[1][1], A[2][2]), mul(A[1][2], A[2][1]));
            B[0][1] = add(mul(-A[1][0], A[2][2]), mul(A[1][2], A[2][0]));
            B[0][2] = sub(mul(A[1][0], A[2][1]), mul(A[1][1], A[2][0]));
            B[1][0] = add(mul(-A[0][1], A[2][2]), mul(A[0][2], A[2][1]));
            B[1][1] = sub(mul(A[0][0], A[2][2]), mul(A[0][2], A[2][0]));
            B[1][2] = add(mul(-A[0][0], A[2][1]), mul(A[0][1], A[2][0]));
            B[2][0] = sub(mul(A[0][1], A[1][2]), mul(A[0][2], A[1][1]));
            B[2][1] = add(mul(-A[0][0], A[1][2]), mul(A[0][2], A[1][0]));
            B[2][2] = sub(mul(A[0][0], A[1][1]), mul(A[0][1], A[1][0]));
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) A[a][b] = div(B[a][b], det);
        }
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) MPHX_T(so.Linv, a, b, s, so.ns) = A[a][b];
}

// K7 "solid pass 1": deformation gradient (:2701-2752), Green-Lagrange strain and 2nd PK stress
// (:2768-2808), and P = F S L^-1 (:2837-2852), all in registers.
template <int DIMS>
__global__ void k_solid_pass1(Solid so, double W0, double W1, double W2, double radius, double cw)
{
    using namespace ex;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const int ns = so.ns;
    const double ui[3] = {so.ux[s], so.uy[s], DIMS == 3 ? so.uz[s] : 0.0};
    double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int len = so.len[s];
    for (int kk = 0; kk < len; ++kk) {
        const size_t k = (size_t)kk * ns + s;
        const int j = so.enbr[k];
        const double d0[3] = {so.d0x[k], so.d0y[k], DIMS == 3 ? so.d0z[k] : 0.0};
        const double uj[3] = {so.ux[j], so.uy[j], DIMS == 3 ? so.uz[j] : 0.0};
        double d[3];
        for (int a = 0; a < DIMS; ++a) d[a] = add(d0[a], sub(uj[a], ui[a])); // :2716
        const double w = so.w[k];
        for (int a = 0; a < DIMS; ++a)
            for (int b = 0; b < DIMS; ++b) G[a][b] = add(G[a][b], mul(mul(w, d[a]), d0[b])); // :2726
    }
    double L[3][3], F[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) L[a][b] = MPHX_T(so.Linv, a, b, s, ns);
    for (int a = 0; a < DIMS; ++a)
        for (int b = 0; b < DIMS; ++b) {
            double sum = 0.0;
            for (int k = 0; k < DIMS; ++k) sum = add(sum, mul(G[a][k], L[k][b])); // :2743
            F[a][b] = sum;
        }
    double E[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, tr = 0.0;
    for (int a = 0; a < DIMS; ++a)
        for (int b = 0; b < DIMS; ++b) {
            double sum = 0.0;
            for (int k = 0; k < DIMS; ++k) sum = add(sum, mul(F[k][a], F[k][b])); // :2780
            E[a][b] = mul(0.5, sub(sum, (a == b ? 1.0 : 0.0)));
            if (a == b) tr = add(tr, E[a][b]);
        }
    const double mu = so.mu[s], lam = so.lam[s];
    for (int a = 0; a < DIMS; ++a)
        for (int b = 0; b < DIMS; ++b) {
            S[a][b] = mul(mul(2.0, mu), E[a][b]); // :2804
            if (a == b) S[a][b] = add(S[a][b], mul(lam, tr));
        }
    for (int a = 0; a < DIMS; ++a)
        for (int b = 0; b < DIMS; ++b) {
            double sum = 0.0;
            for (int k = 0; k < DIMS; ++k)
                for (int l = 0; l < DIMS; ++l) sum = add(sum, mul(mul(F[a][k], S[k][l]), L[l][b])); // :2847
            MPHX_T(so.Pk, a, b, s, ns) = sum;
            MPHX_T(so.Fm, a, b, s, ns) = F[a][b];
            MPHX_T(so.E, a, b, s, ns) = E[a][b];
            MPHX_T(so.S, a, b, s, ns) = S[a][b];
        }
}

// K8 "solid pass 2": the reference scatters  v_i += w P_i x0_ij /rho_i dt,  v_j -= (same)/rho_j dt
// serially / with atomics (:2855-2887); here every particle GATHERS, in the reference's serial
// order, the terms of rows j<s that list s, then its own row, then rows j>s (transposed list), so
// the result is deterministic, atomic-free and equal to the reference's CPU bits.
// Then updateElasticPosition (:1916-2081) incl. the clamp modules and quirk Q1.
template <int DIMS>
__global__ void k_solid_pass2(Solid so, double W0, double W1, double W2, double radius, double cw, double edt,
                              int module, int double_update, const double *__restrict__ inv_density)
{
    using namespace ex;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const int ns = so.ns;
    const double xi0 = so.x0[s], yi0 = so.y0[s], zi0 = so.z0[s];
    const double ir = inv_density[so.type[s]];
    double v[3] = {so.vx[s], so.vy[s], so.vz[s]};
    auto scattered_from = [&](int kk, int j) { // row j lists s:  v_s -= invRho_s * (w P_j x0_js) * dt
        const size_t k = (size_t)kk * ns + s;
        const double d0[3] = {so.rd0x[k], so.rd0y[k], DIMS == 3 ? so.rd0z[k] : 0.0};
        const double w = so.rw[k];
        for (int a = 0; a < DIMS; ++a) {
            double f = 0.0;
            for (int b = 0; b < DIMS; ++b) f = add(f, mul(MPHX_T(so.Pk, a, b, j, ns), d0[b]));
            f = mul(f, w);
            v[a] = sub(v[a], mul(mul(ir, f), edt)); // :2885
        }
    };
    int kr = 0;
    const int rlen = so.rlen[s];
    for (; kr < rlen; ++kr) {
        const int j = so.ernbr[(size_t)kr * ns + s];
        if (j >= s) break;
        scattered_from(kr, j);
    }
    {
        double Pi[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Pi[a][b] = (a < DIMS && b < DIMS) ? MPHX_T(so.Pk, a, b, s, ns) : 0.0;
        const int len = so.len[s];
        for (int kk = 0; kk < len; ++kk) { // own row: v_s += invRho_s * (w P_s x0_sj) * dt
            const size_t q = (size_t)kk * ns + s;
            const double d0[3] = {so.d0x[q], so.d0y[q], DIMS == 3 ? so.d0z[q] : 0.0};
            const double w = so.w[q];
            for (int a = 0; a < DIMS; ++a) {
                double f = 0.0;
                for (int b = 0; b < DIMS; ++b) f = add(f, mul(Pi[a][b], d0[b]));
                f = mul(f, w);
                v[a] = add(v[a], mul(mul(ir, f), edt)); // :2883
            }
        }
    }
    for (; kr < rlen; ++kr) scattered_from(kr, so.ernbr[(size_t)kr * ns + s]);
    double x[3] = {so.x[s], so.y[s], so.z[s]};
    // Acceleration of solids is 0 (:2892): v += 0*dt leaves v unchanged
    if (module != 0) {
        const bool clamped = (module == 1) ? (xi0 < 0.001) : (yi0 < 0.002); // :1919 / :1968
        if (clamped) {
            x[0] = xi0; x[1] = yi0; x[2] = zi0;
            v[0] = v[1] = v[2] = 0.0;
            so.fx[s] = 0.0; so.fy[s] = 0.0; so.fz[s] = 0.0;
        } else {
            for (int a = 0; a < 3; ++a) x[a] = add(x[a], mul(v[a], edt));
        }
        if (double_update) // Q1 (:2070-2079)
            for (int a = 0; a < 3; ++a) x[a] = add(x[a], mul(v[a], edt));
    } else {
        for (int a = 0; a < 3; ++a) x[a] = add(x[a], mul(v[a], edt));
    }
    so.x[s] = x[0]; so.y[s] = x[1]; so.z[s] = x[2];
    so.vx[s] = v[0]; so.vy[s] = v[1]; so.vz[s] = v[2];
    so.ux[s] = minimg_exact(x[0], xi0, W0); so.uy[s] = minimg_exact(x[1], yi0, W1); so.uz[s] = minimg_exact(x[2], zi0, W2);
}

// static pair data of the reference configuration (once, after the lists are known)
template <int DIMS>
__global__ void k_solid_pairs(Solid so, double W0, double W1, double W2, double radius, double cw)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const double xi0 = so.x0[s], yi0 = so.y0[s], zi0 = so.z0[s];
    const int ns = so.ns;
    int kk = 0;
    for (int c = so.off[s]; c < so.off[s + 1]; ++c, ++kk) {
        const int j = so.nbr[c];
        const size_t k = (size_t)kk * ns + s;
        const double a = minimg_exact(so.x0[j], xi0, W0), b = minimg_exact(so.y0[j], yi0, W1);
        const double cc = DIMS == 3 ? minimg_exact(so.z0[j], zi0, W2) : 0.0;
        so.enbr[k] = j;
        so.d0x[k] = a; so.d0y[k] = b; so.d0z[k] = cc;
        so.w[k] = tl_weight<DIMS>(a, b, cc, radius, cw);
    }
    so.len[s] = kk;
    kk = 0;
    for (int c = so.roff[s]; c < so.roff[s + 1]; ++c, ++kk) {
        const int j = so.rnbr[c]; // row j lists s: x0_js = Mod(x0_s - x0_j ...) as row j computes it
        const size_t k = (size_t)kk * ns + s;
        const double a = minimg_exact(xi0, so.x0[j], W0), b = minimg_exact(yi0, so.y0[j], W1);
        const double cc = DIMS == 3 ? minimg_exact(zi0, so.z0[j], W2) : 0.0;
        so.ernbr[k] = j;
        so.rd0x[k] = a; so.rd0y[k] = b; so.rd0z[k] = cc;
        so.rw[k] = tl_weight<DIMS>(a, b, cc, radius, cw);
    }
    so.rlen[s] = kk;
}

// ------------------------------------------------------------------------------------------------
// upload / download helpers (AoS original order <-> sorted SoA)
__global__ void k_upload_split(int n, const int *__restrict__ type, const double *__restrict__ x3,
                               const double *__restrict__ v3, Particles p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p.x[i] = x3[3 * (size_t)i]; p.y[i] = x3[3 * (size_t)i + 1]; p.z[i] = x3[3 * (size_t)i + 2];
    p.vx[i] = v3[3 * (size_t)i]; p.vy[i] = v3[3 * (size_t)i + 1]; p.vz[i] = v3[3 * (size_t)i + 2];
    p.type[i] = type[i]; p.id[i] = i; p.key[i] = 0;
}
__global__ void k_solid_upload(Solid so, const int *__restrict__ type, const double *__restrict__ x3,
                               const double *__restrict__ x03, const double *__restrict__ v3)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const size_t i = (size_t)(so.sb + s);
    so.x[s] = x3[3 * i]; so.y[s] = x3[3 * i + 1]; so.z[s] = x3[3 * i + 2];
    so.x0[s] = x03[3 * i]; so.y0[s] = x03[3 * i + 1]; so.z0[s] = x03[3 * i + 2];
    so.vx[s] = v3[3 * i]; so.vy[s] = v3[3 * i + 1]; so.vz[s] = v3[3 * i + 2];
    so.fx[s] = so.fy[s] = so.fz[s] = 0.0;
    so.type[s] = type[i];
}
// host state (original order AoS) -> current sorted slots (+ the solid arrays)
__global__ void k_upload_state(int n, Particles p, Solid sol, const double *__restrict__ x3, const double *__restrict__ v3)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int id = p.id[q];
    const size_t o = 3 * (size_t)id;
    const double x = x3[o], y = x3[o + 1], z = x3[o + 2], vx = v3[o], vy = v3[o + 1], vz = v3[o + 2];
    p.x[q] = x; p.y[q] = y; p.z[q] = z; p.vx[q] = vx; p.vy[q] = vy; p.vz[q] = vz;
    if (is_structure_type(p.type[q])) {
        const int s = id - sol.sb;
        sol.x[s] = x; sol.y[s] = y; sol.z[s] = z; sol.vx[s] = vx; sol.vy[s] = vy; sol.vz[s] = vz;
    }
}
__global__ void k_split_vec3(int n, const double *__restrict__ a3, double *__restrict__ x, double *__restrict__ y,
                             double *__restrict__ z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = a3[3 * (size_t)i]; y[i] = a3[3 * (size_t)i + 1]; z[i] = a3[3 * (size_t)i + 2];
}
__global__ void k_restore_by_id(int n, const int *__restrict__ id, const double *__restrict__ sx, const double *__restrict__ sy,
                                const double *__restrict__ sz, double *__restrict__ x, double *__restrict__ y,
                                double *__restrict__ z)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int i = id[q];
    x[q] = sx[i]; y[q] = sy[i]; z[q] = sz[i];
}
// sorted SoA vec3 -> original-order AoS
__global__ void k_gather_vec3(int n, const int *__restrict__ id, const double *__restrict__ a, const double *__restrict__ b,
                              const double *__restrict__ c, double *__restrict__ out3)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const size_t o = 3 * (size_t)id[q];
    out3[o] = a[q]; out3[o + 1] = b[q]; out3[o + 2] = c[q];
}
__global__ void k_gather_scalar(int n, const int *__restrict__ id, const double *__restrict__ a, double *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[id[q]] = a[q];
}
__global__ void k_gather_int(int n, const int *__restrict__ id, const int *__restrict__ a, int *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[id[q]] = a[q];
}
// solid arrays -> original-order AoS (overrides the stale sorted copies)
__global__ void k_solid_vec3_to_orig(Solid so, const double *__restrict__ a, const double *__restrict__ b,
                                     const double *__restrict__ c, double *__restrict__ out3)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const size_t o = 3 * (size_t)(so.sb + s);
    out3[o] = a[s]; out3[o + 1] = b[s]; out3[o + 2] = c[s];
}
__global__ void k_solid_scalar_to_orig(Solid so, const double *__restrict__ a, double *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    out[so.sb + s] = a[s];
}
__global__ void k_solid_tensor_to_orig(Solid so, const double *__restrict__ M, double *__restrict__ out9)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    const size_t o = 9 * (size_t)(so.sb + s);
    for (int k = 0; k < 9; ++k) out9[o + k] = M[(size_t)k * so.ns + s];
}
__global__ void k_solid_rowlen_to_orig(Solid so, int *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    out[so.sb + s] = so.off[s + 1] - so.off[s];
}

} // namespace mphx
