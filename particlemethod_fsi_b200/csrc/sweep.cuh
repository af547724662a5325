// sweep.cuh -- pass 1 (K5) and pass 2 (K6) of the explicit step, and the one-particle candidate filter.
//
// Normal operation (see also sweep_pair.cuh):
//   k_filter2 / k_filter  test every candidate of a particle's stencil with a CONSERVATIVE single-precision
//       distance filter on `pf` (positions in bucket units, pair-interleaved 24-byte records, packed f32x2
//       arithmetic) and write the survivors into a per-step ELL candidate list (entry k of particle i at
//       nbr[k*cap + i]: coalesced).  No fp64, no shared-memory queue, few registers: full occupancy.
//   k_pass1_v3<.., LIST=true>, k_pass2_v3<.., LIST=true>  traverse that list -- both passes run in the same
//       step on the same positions -- two entries per trip: exact double-precision separation from the
//       32-byte gather records (x,y,z,vx | vy,vz,P,type; one 256-bit load each), the reference's exact
//       cut-off tests, kernel weights and pair terms as straight-line (masked) code.
//
// The filter only has to be a superset of the exact predicates: the float coordinates carry an error
// <= 2 ulp_f32(max bucket coordinate) each, which the host turns into a margin on the squared cut-off
// (filter_radius2 in mphx.cu).  Physics is decided in fp64.  (A particle always passes its own filter
// test; the passes reject j == i.)
//
// Fall-back (LIST=false): the FUSED sweep -- one thread per particle walks its stencil in batches of
// columns, phase A (the same filter, survivors pushed to a per-thread queue in shared memory: slot-major
// layout, bank = lane, conflict-free) and phase B (drain the queue: the same pair terms) per batch.  It
// finishes the particles whose list overflowed (count > L) and serves MPHX_LIST_CAP=0; with a list it is
// launched as a small persistent grid that returns at once when the step's overflow flag is clear.
//
// The reference procedures these kernels replace are listed at the top of kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace mphx {

constexpr int kSweepThreads = 128;
// minimum resident blocks per SM the register allocation of each kernel is tuned for
#ifndef MPHX_FILTER_MINB
#define MPHX_FILTER_MINB 10
#endif
#ifndef MPHX_P1_MINB
#define MPHX_P1_MINB 8
#endif
#ifndef MPHX_P2_MINB
#define MPHX_P2_MINB 6
#endif
// the surface-tension instantiations carry ~50 more live values per pair (DensityA, GravityCenter, the
// diffuse-interface terms): fewer resident blocks instead of spills
#ifndef MPHX_P1ST_MINB
#define MPHX_P1ST_MINB 5
#endif
#ifndef MPHX_P2ST_MINB
#define MPHX_P2ST_MINB 3
#endif
// straight-line (masked) pair bodies in pass 1 / pass 2 instead of branches
#ifndef MPHX_BRANCHFREE
#define MPHX_BRANCHFREE 1
#endif
constexpr int kQueueCap = 40; // queue slots per thread (rows of 128 uint); one spare row absorbs masked stores

struct SweepShared {
    unsigned q[kQueueCap + 1][kSweepThreads];
    int sdx[kMaxStencil], sdy[kMaxStencil], sh[kMaxStencil];
    int dlo[kMaxStencil], dhi[kMaxStencil]; // bucket offsets of a column's run ends relative to the own key
};

__device__ __forceinline__ void load_stencil(SweepShared &sm, const GridDesc &g)
{
    for (int e = threadIdx.x; e < g.nsten; e += blockDim.x) {
        const int dx = g.sdx[e], dy = g.sdy[e], h = g.sh[e];
        sm.sdx[e] = dx; sm.sdy[e] = dy; sm.sh[e] = h;
        const int d = (g.dim == 3) ? (dx * g.ny + dy) * g.nz : dx * g.ny;
        sm.dlo[e] = d - h;
        sm.dhi[e] = d + h + 1;
    }
}

// 1/sqrt(a) for a > 0: hardware seed (MUFU.RSQ64H) + two Newton steps (error ~1 ulp).
__device__ __forceinline__ double rsqrt_nr(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}

// 256-bit loads (LDG.E.256 on sm_100a): one request per 32-byte record.
__device__ __forceinline__ Rec ld_rec_nc(const Rec *p)
{
    Rec r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
// coherent variants: pass 1 reads (vy, vz) from records whose PressureP slot other threads are writing.
// FULL = false fetches only the first half (vy, vz); c, d are then unspecified.
template <bool FULL> __device__ __forceinline__ Rec ld_rec(const Rec *p)
{
    Rec r;
    if (FULL) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    else { asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.a), "=d"(r.b) : "l"(p)); r.c = 0.0; r.d = 0.0; }
    return r;
}
__device__ __forceinline__ PfPair ld_pf_nc(const PfPair *p)
{
    const float2 *q = reinterpret_cast<const float2 *>(p);
    PfPair r;
    r.x = __ldg(q); r.y = __ldg(q + 1); r.z = __ldg(q + 2);
    return r;
}

struct Bucket3 { int cx, cy, cr, nr; };
template <int DIM> __device__ __forceinline__ Bucket3 split_key(const GridDesc &g, int key)
{
    Bucket3 b;
    if (DIM == 3) { b.cr = key % g.nz; const int t = key / g.nz; b.cy = t % g.ny; b.cx = t / g.ny; b.nr = g.nz; }
    else { b.cr = key % g.ny; b.cx = key / g.ny; b.cy = 0; b.nr = g.ny; }
    return b;
}
// does any part of the stencil of this bucket cross the periodic box?
template <int DIM> __device__ __forceinline__ bool stencil_wraps(const GridDesc &g, const Bucket3 &b)
{
    const int R = g.range;
    bool w = (b.cx - R < 0) || (b.cx + R >= g.nx) || (b.cr - R < 0) || (b.cr + R >= b.nr);
    if (DIM == 3) w = w || (b.cy - R < 0) || (b.cy + R >= g.ny);
    return w;
}

struct MinImage {
    double W0, W1, W2, h0, h1, h2;
    __device__ __forceinline__ explicit MinImage(const GridDesc &g)
        : W0(g.W[0]), W1(g.W[1]), W2(g.W[2]), h0(0.5 * g.W[0]), h1(0.5 * g.W[1]), h2(0.5 * g.W[2]) {}
    __device__ __forceinline__ void apply(double &dx, double &dy, double &dz) const
    {
        dx = dx > h0 ? dx - W0 : (dx < -h0 ? dx + W0 : dx);
        dy = dy > h1 ? dy - W1 : (dy < -h1 ? dy + W1 : dy);
        dz = dz > h2 ? dz - W2 : (dz < -h2 ? dz + W2 : dz);
    }
};

// Walks the stencil of particle i in batches of `batch` columns; calls hit(j, dx, dy, dz, r2, vxj) in
// phase B for every candidate that passed the fp32 filter (INCLUDING j == i, which the caller
// rejects).  dx,dy,dz,r2 are the exact fp64 separation (minimum image) of j from i; vxj rides along
// in the same 32-byte gather record.
//
// Phase A works on PAIRS of candidates: the filter coordinates are stored pair-interleaved
// (PfPair = x0 x1 y0 y1 z0 z1 t0 t1, one 256-bit load), so the three subtractions and the three
// multiply-adds of the distance test are Blackwell's packed f32x2 instructions (FADD2/FMUL2/FFMA2:
// two candidates per issue slot).  Pushes are branch-free: the candidate index is always stored at
// the queue top and the top only advances when the test passed.
template <int DIM, class Hit>
__device__ __forceinline__ void sweep(SweepShared &sm, const GridDesc &g, const int *__restrict__ cellStart,
                                      const PfPair *__restrict__ pf, const Rec *__restrict__ ra, int i, bool active,
                                      int key, double xi, double yi, double zi, float filt2, int batch, Hit &&hit)
{
    const int tid = threadIdx.x;
    float fxi = 0.f, fyi = 0.f, fzi = 0.f;
    if (active) {
        const float *f = reinterpret_cast<const float *>(pf + (i >> 1)) + (i & 1);
        fxi = f[0]; fyi = f[2]; fzi = f[4];
    }
    const Bucket3 b = split_key<DIM>(g, key);
    const int cx = b.cx, cy = b.cy, cr = b.cr, nr = b.nr;
    const bool wraps = stencil_wraps<DIM>(g, b);
    const bool warp_wraps = __any_sync(0xffffffffu, active && wraps);
    const MinImage mi(g);

    unsigned *const q0 = &sm.q[0][tid];
    unsigned *qp = q0; // queue top
    auto queued = [&]() { return (int)(qp - q0) / kSweepThreads; };
    auto drain = [&]() {
        const int cnt = queued();
        // two queue entries per trip: both record loads are issued before either is used
        for (int s = 0; s < cnt; s += 2) {
            const bool two = s + 1 < cnt;
            const int j0 = (int)q0[s * kSweepThreads];
            const int j1 = two ? (int)q0[(s + 1) * kSweepThreads] : j0;
            const Rec a0 = ld_rec_nc(ra + j0);
            const Rec a1 = ld_rec_nc(ra + j1);
            double dx0 = a0.a - xi, dy0 = a0.b - yi, dz0 = a0.c - zi;
            double dx1 = a1.a - xi, dy1 = a1.b - yi, dz1 = a1.c - zi;
            if (warp_wraps) { mi.apply(dx0, dy0, dz0); mi.apply(dx1, dy1, dz1); }
            hit(j0, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0, a0.d);
            if (two) hit(j1, dx1, dy1, dz1, dx1 * dx1 + dy1 * dy1 + dz1 * dz1, a1.d);
        }
        qp = q0;
    };

    // phase A over one contiguous run [jb, je) of candidates (the caller guarantees queue room for
    // je - jb + 1 entries).  Pairs (2k, 2k+1) are loaded whole; the elements outside [jb, je) are
    // masked by the range test.  Two pair loads are in flight per trip (pf is padded by two pairs).
    auto scan_run = [&](int jb, int je, float fx, float fy, float fz) {
        const float2 nx = make_float2(-fx, -fx), ny = make_float2(-fy, -fy), nz = make_float2(-fz, -fz);
        const int len = je - jb, lenm1 = len - 1;
        int j0 = jb & ~1;
        int t = j0 - jb; // -1 or 0: position of the pair's first element in the run
        const PfPair *pp = pf + (j0 >> 1);
        for (; t < len; t += 4, j0 += 4, pp += 2) {
            const PfPair fa = ld_pf_nc(pp), fb = ld_pf_nc(pp + 1);
#define MPHX_TEST2(f, tt, jj)                                                                      \
    {                                                                                              \
        const float2 ddx = __fadd2_rn(f.x, nx), ddy = __fadd2_rn(f.y, ny), ddz = __fadd2_rn(f.z, nz); \
        float2 d2 = __fmul2_rn(ddx, ddx);                                                          \
        d2 = __ffma2_rn(ddy, ddy, d2);                                                             \
        d2 = __ffma2_rn(ddz, ddz, d2);                                                             \
        const bool ok0 = (unsigned)(tt) < (unsigned)len && d2.x <= filt2;                          \
        const bool ok1 = (tt) < lenm1 && d2.y <= filt2;                                            \
        *qp = (unsigned)(jj); qp += ok0 ? kSweepThreads : 0;                                       \
        *qp = (unsigned)(jj) + 1u; qp += ok1 ? kSweepThreads : 0;                                  \
    }
            MPHX_TEST2(fa, t, j0) MPHX_TEST2(fb, t + 2, j0 + 2)
#undef MPHX_TEST2
        }
    };
    auto scan_checked = [&](int jb, int je, float fx, float fy, float fz) {
        while (jb < je) { // make room first; a run longer than the queue is scanned in pieces
            const int room = kQueueCap - queued();
            if (room < 2) { drain(); continue; }
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
            const int *cs = cellStart + key;
            int jbn = cs[sm.dlo[e0]], jen = cs[sm.dhi[e0]];
            for (int e = e0; e < e1; ++e) {
                const int jb = jbn, je = jen;
                if (e + 1 < e1) { jbn = cs[sm.dlo[e + 1]]; jen = cs[sm.dhi[e + 1]]; }
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
                    int a, bb;
                    float shf = 0.f;
                    if (seg == 0) { a = lo < 0 ? 0 : lo; bb = hi >= nr ? nr - 1 : hi; }
                    else if (seg == 1) { if (lo >= 0) continue; a = lo + nr; bb = nr - 1; shf = (float)nr; }
                    else { if (hi < nr) break; a = 0; bb = hi - nr; shf = -(float)nr; }
                    const float fyy = (DIM == 2) ? fy + shf : fy;
                    const float fzz = (DIM == 3) ? fz + shf : fz;
                    scan_checked(cellStart[base + a], cellStart[base + bb + 1], fx, fyy, fzz);
                }
            }
        }
        drain();
    }
}

// Restriction of a pass-2 launch: only the sorted slots listed (the solids), or everything but the solids.
// Lets the solid sub-steps start -- on a second stream -- as soon as the solids' share of pass 2 is done.
struct Subset {
    const int *slots; // nullptr: thread t handles particle t
    int count;
    int skip_solids;
};

// Candidate list of one step: written by k_filter, traversed by pass 1 and pass 2 (same step, same
// positions).  It holds the survivors of the conservative fp32 filter (a superset of every exact
// cut-off test, possibly including the particle itself); the physics kernels apply the exact fp64
// tests.  ELL layout: entry k of particle i at nbr[k*cap + i] (coalesced across a warp).
struct PairList {
    int *nbr;    // [L + 1][cap]; row L is a parking row for the stores of overflowed lists
    int *count;  // [cap] entries of particle i; L + 1 = overflowed (that particle is swept instead)
    int *flags;  // [0] != 0: some list overflowed in this step
    int cap, L;
    const unsigned char *skip; // pass 1 only: skip[i] != 0 = particle i is handled by the staged kernel (brick.cuh)
};

// slab mode: ghosts and parked solids are only ever neighbours; a (replicated) solid is evaluated by the slab
// that owned its column when the list was built
__device__ __forceinline__ bool particle_active(const GridDesc &g, int i, int n, int tflag, int key)
{
    bool active = i < n && key < g.ncells && !(tflag & kGhost);
    if (g.slab && is_structure_type(tflag)) active = i < n && !(tflag & kGhost) && (tflag & kSolidOwned);
    return active;
}

// K5a "filter": phase A alone, for every particle at once.  No fp64, no shared-memory queue, few
// registers: the kernel runs at full occupancy, which hides the latency of the candidate loads.
// Candidates are tested in pairs with the packed f32x2 instructions (see sweep()).
template <int DIM>
__global__ void __launch_bounds__(kSweepThreads, MPHX_FILTER_MINB)
k_filter(const Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, PairList pl)
{
    if (!ctl->rebuild) return; // the list of an earlier step is still a superset of every cut-off set
    const int n = ctl->n;
    const float filt2 = ctl->filt2;
    __shared__ int s_dlo[kMaxStencil], s_dhi[kMaxStencil], s_sdx[kMaxStencil], s_sdy[kMaxStencil], s_sh[kMaxStencil];
    for (int e = threadIdx.x; e < g.nsten; e += blockDim.x) {
        const int dx = g.sdx[e], dy = g.sdy[e], h = g.sh[e];
        const int d = (DIM == 3) ? (dx * g.ny + dy) * g.nz : dx * g.ny;
        s_dlo[e] = d - h; s_dhi[e] = d + h + 1; s_sdx[e] = dx; s_sdy[e] = dy; s_sh[e] = h;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int tflag = p.type[i], keyi = p.key[i];
    const bool active = particle_active(g, i, n, tflag, keyi);
    if (!active) { pl.count[i] = 0; return; }
    const PfPair *__restrict__ pf = p.pf;
    const float *fo = reinterpret_cast<const float *>(pf + (i >> 1)) + (i & 1);
    const float fxi = fo[0], fyi = fo[2], fzi = fo[4];
    const Bucket3 b = split_key<DIM>(g, keyi);
    const bool wraps = stencil_wraps<DIM>(g, b);
    // list top as a 32-bit element offset into nbr (host guarantees (L+1)*cap < 2^32); a push past
    // row L-1 lands on the parking row (offset `park`) and stays there: no capacity branch.
    int *__restrict__ nbr = pl.nbr;
    unsigned stride = (unsigned)pl.cap;
    unsigned park = (unsigned)pl.L * stride + (unsigned)i;
    float filt = filt2;
    // keep the loop invariants in registers (the compiler otherwise re-reads them from the constant
    // bank under every predicated push)
    asm volatile("" : "+l"(nbr), "+r"(stride), "+r"(park), "+f"(filt));
    unsigned off = (unsigned)i;

    auto scan_run = [&](int jb, int je, float fx, float fy, float fz) {
        const int len = je - jb, lenm1 = len - 1;
        const float2 nx = make_float2(-fx, -fx), ny = make_float2(-fy, -fy), nz = make_float2(-fz, -fz);
        int j0 = jb & ~1;
        int t = j0 - jb; // -1 or 0: position of the pair's first element in the run
        const PfPair *pp = pf + (j0 >> 1);
        for (; t < len; t += 4, j0 += 4, pp += 2) {
            const PfPair fa = ld_pf_nc(pp);
            PfPair fb = fa;
            if (t + 2 < len) fb = ld_pf_nc(pp + 1); // (its tests are masked by the range check otherwise)
#define MPHX_PUSH(ok, jj) if (ok) { nbr[off] = (jj); off = min(off + stride, park); }
#define MPHX_TEST2(f, tt, jj)                                                                      \
    {                                                                                              \
        const float2 ddx = __fadd2_rn(f.x, nx), ddy = __fadd2_rn(f.y, ny), ddz = __fadd2_rn(f.z, nz); \
        float2 d2 = __fmul2_rn(ddx, ddx);                                                          \
        d2 = __ffma2_rn(ddy, ddy, d2);                                                             \
        d2 = __ffma2_rn(ddz, ddz, d2);                                                             \
        MPHX_PUSH((unsigned)(tt) < (unsigned)len && d2.x <= filt, (jj))                            \
        MPHX_PUSH((tt) < lenm1 && d2.y <= filt, (jj) + 1)                                          \
    }
            MPHX_TEST2(fa, t, j0) MPHX_TEST2(fb, t + 2, j0 + 2)
#undef MPHX_TEST2
#undef MPHX_PUSH
        }
    };

    const int nsten = g.nsten;
    if (!wraps) {
        // stencil inside the box: one run per column; the bounds of the next column are fetched
        // while the current run is scanned
        int jbn = __ldg(cellStart + (keyi + s_dlo[0])), jen = __ldg(cellStart + (keyi + s_dhi[0]));
        for (int e = 0; e < nsten; ++e) {
            const int jb = jbn, je = jen;
            const int en = (e + 1 < nsten) ? e + 1 : e; // (the last trip re-reads its own bounds: no branch)
            jbn = __ldg(cellStart + (keyi + s_dlo[en])); jen = __ldg(cellStart + (keyi + s_dhi[en]));
            scan_run(jb, je, fxi, fyi, fzi);
        }
    } else {
        const int cx = b.cx, cy = b.cy, cr = b.cr, nr = b.nr;
        for (int e = 0; e < nsten; ++e) {
            int ccx = cx + s_sdx[e];
            float fx = fxi, fy = fyi, fz = fzi;
            if (ccx < 0) { ccx += g.nx; fx += (float)g.nx; }
            else if (ccx >= g.nx) { ccx -= g.nx; fx -= (float)g.nx; }
            int base;
            if (DIM == 3) {
                int ccy = cy + s_sdy[e];
                if (ccy < 0) { ccy += g.ny; fy += (float)g.ny; }
                else if (ccy >= g.ny) { ccy -= g.ny; fy -= (float)g.ny; }
                base = (ccx * g.ny + ccy) * g.nz;
            } else {
                base = ccx * g.ny;
            }
            const int h = s_sh[e];
            const int lo = cr - h, hi = cr + h;
#pragma unroll 1
            for (int seg = 0; seg < 3; ++seg) { // in-range part, then the wrapped images
                int a, bb;
                float shf = 0.f;
                if (seg == 0) { a = lo < 0 ? 0 : lo; bb = hi >= nr ? nr - 1 : hi; }
                else if (seg == 1) { if (lo >= 0) continue; a = lo + nr; bb = nr - 1; shf = (float)nr; }
                else { if (hi < nr) break; a = 0; bb = hi - nr; shf = -(float)nr; }
                const float fyy = (DIM == 2) ? fy + shf : fy;
                const float fzz = (DIM == 3) ? fz + shf : fz;
                scan_run(cellStart[base + a], cellStart[base + bb + 1], fx, fyy, fzz);
            }
        }
    }
    if (off == park) { pl.count[i] = pl.L + 1; atomicOr(pl.flags, 1); } // (a list of exactly L entries counts as overflowed)
    else pl.count[i] = (int)((off - (unsigned)i) / stride);
}

// K5 "pass 1": VolStrainP, DivergenceP -> PressureP (+ DensityA, GravityCenter, PressureA when any
// surface tension is set).  All particle classes (:2320, :2349).
//   LIST = true : candidates come from k_filter's list; particles whose list overflowed are skipped.
//   LIST = false: fused stencil sweep (phase A + phase B); with pl.count set only the overflowed
//                 particles are processed (the block exits at once if it has none), else all.
template <int DIM, bool ST, bool LIST>
__device__ __forceinline__ void
pass1_block(int vblock, int n, Particles p, const int *__restrict__ cellStart, const GridDesc &g, const Phys &ph, float filt2, int batch,
           double *__restrict__ P, double *__restrict__ volStrain, double *__restrict__ divP,
           double *__restrict__ densA, double *__restrict__ gcx, double *__restrict__ gcy, double *__restrict__ gcz,
           double *__restrict__ PA, PairList pl)
{
    const int i0 = vblock * blockDim.x + threadIdx.x;
    const int i = i0 < n ? i0 : n - 1;
    int mycount = 0;
    bool mine = i0 < n;
    if (pl.count) {
        mycount = pl.count[i];
        mine = mine && (LIST ? mycount <= pl.L : mycount > pl.L);
        if (LIST && pl.skip && pl.skip[i]) mine = false; // done from shared memory by k_brick_pass1
        if (!LIST && !__syncthreads_or(mine ? 1 : 0)) return; // no overflowed particle in this block
    }
    const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const int tflag = p.type[i], ti = real_type(tflag), keyi = p.key[i];
    const bool solid_i = is_structure_type(ti);
    // slab mode: ghosts and parked solids are only ever neighbours; a solid is evaluated by the slab
    // that owns its current column
    const bool active = particle_active(g, i0, n, tflag, keyi);
    const double rp2 = ph.rp2, irp = ph.irp, ra2 = ph.ra2, ira = ph.ira;
    double nP = 0.0, dv = 0.0, nA = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0; // nP, dv without their constant factors
    const Rec *__restrict__ RA = p.ra;
    const Rec *__restrict__ RB = p.rb;
    // pair terms; rb = (vy, vz, -, type bits) of j
    auto pair = [&](int j, double dx, double dy, double dz, double r2, double vxj, const Rec rb) {
        const bool other = j != i;
#if MPHX_BRANCHFREE
        { // straight-line form: a masked pair contributes exact zeros, and the FP64 chains of the two
          // pairs of a trip can be interleaved by the scheduler (no reconvergence regions in between)
            const bool in = other && r2 <= rp2; // :2333, :2362
            const double r2s = in ? r2 : 1.0;
            const double rinv = rsqrt_nr(r2s);
            const double q = in ? 1.0 - (r2s * rinv) * irp : 0.0;
            nP += q * q;
            const double ux = vxj - vxi, uy = rb.a - vyi, uz = rb.b - vzi;
            dv -= (ux * dx + uy * dy + uz * dz) * rinv * q;
        }
#else
        if (other && r2 <= rp2) { // :2333, :2362
            const double rinv = rsqrt_nr(r2);
            const double q = 1.0 - (r2 * rinv) * irp;
            nP += q * q;
            const double ux = vxj - vxi, uy = rb.a - vyi, uz = rb.b - vzi;
            dv -= (ux * dx + uy * dy + uz * dz) * rinv * q;
        }
#endif
        if (ST && other && !solid_i && r2 <= ra2) { // :2162, :2195
            const double r = sqrt(r2);
            const double qa = r * ira;
            const double ratio = ph.ratio[ti][(int)__double_as_longlong(rb.d)];
            nA += ratio * (ph.cwa * qa * (1.0 - qa) * (1.0 - qa));
            const double wgv = ratio * (ph.cwg * ((1.0 - qa) * (1.0 - qa))) / ph.r2g * ph.rg;
            g0 += dx * wgv; g1 += dy * wgv; g2 += dz * wgv;
        }
    };
    if (LIST) {
        const Bucket3 b = split_key<DIM>(g, active ? keyi : 0);
        const bool warp_wraps = __any_sync(0xffffffffu, active && mine && stencil_wraps<DIM>(g, b));
        const MinImage mi(g);
        const int cnt = (active && mine) ? mycount : 0;
        const int *lp = pl.nbr + i;
        const size_t ls = (size_t)pl.cap;
        int k = 0;
        int j0 = 0, j1 = 0;
        if (cnt > 0) j0 = __ldcs(lp);
        if (cnt > 1) j1 = __ldcs(lp + ls);
        for (; k + 1 < cnt; k += 2) { // two pairs per trip; the next two list entries are fetched first
            const int a_j = j0, b_j = j1;
            if (k + 2 < cnt) j0 = __ldcs(lp + (size_t)(k + 2) * ls);
            if (k + 3 < cnt) j1 = __ldcs(lp + (size_t)(k + 3) * ls);
            const Rec a0 = ld_rec_nc(RA + a_j), a1 = ld_rec_nc(RA + b_j);
            const Rec c0 = ld_rec<ST>(RB + a_j), c1 = ld_rec<ST>(RB + b_j);
            double dx0 = a0.a - xi, dy0 = a0.b - yi, dz0 = a0.c - zi;
            double dx1 = a1.a - xi, dy1 = a1.b - yi, dz1 = a1.c - zi;
            if (warp_wraps) { mi.apply(dx0, dy0, dz0); mi.apply(dx1, dy1, dz1); }
            pair(a_j, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0, a0.d, c0);
            pair(b_j, dx1, dy1, dz1, dx1 * dx1 + dy1 * dy1 + dz1 * dz1, a1.d, c1);
        }
        if (k < cnt) {
            const Rec a0 = ld_rec_nc(RA + j0), c0 = ld_rec<ST>(RB + j0);
            double dx0 = a0.a - xi, dy0 = a0.b - yi, dz0 = a0.c - zi;
            if (warp_wraps) mi.apply(dx0, dy0, dz0);
            pair(j0, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0, a0.d, c0);
        }
    } else {
        __shared__ SweepShared sm;
        load_stencil(sm, g);
        __syncthreads();
        const bool go = active && mine;
        sweep<DIM>(sm, g, cellStart, p.pf, p.ra, i, go, go ? keyi : 0, xi, yi, zi, filt2, batch,
            [&](int j, double dx, double dy, double dz, double r2, double vxj) {
                if (r2 <= rp2 || (ST && r2 <= ra2)) pair(j, dx, dy, dz, r2, vxj, ld_rec<ST>(RB + j));
            });
    }
    if (!mine || !active) return;
    nP *= ph.cwp;
    dv *= ph.cdp;
    const double vs = nP - ph.n0p;                        // :2339
    const double kappa = (vs < 0.0) ? 0.0 : ph.bulk[ti]; // :2112-2113
    double pr = -ph.lambda[ti] * dv;                      // :2388
    if (vs > 0.0) pr += kappa * vs;                       // :2389-2391
    P[i] = pr; volStrain[i] = vs; divP[i] = dv;
    p.rb[i].c = pr; // the gather record pass 2 reads
    if (ST) {
        const double da = solid_i ? 0.0 : nA;
        densA[i] = da;
        gcx[i] = solid_i ? 0.0 : g0; gcy[i] = solid_i ? 0.0 : g1; gcz[i] = solid_i ? 0.0 : g2;
        double pa = ph.cofa[ti] * (da - ph.n0a) / ph.l0; // :2219
        if (ph.n0a <= da) pa = 0.0;
        PA[i] = pa;
    }
}

// kernel wrappers.  The list variants run one block per 128 particles.  The fused-sweep variants are
// normally fall-backs that find nothing to do: they are launched as a small persistent grid that first
// reads the step's overflow flag and otherwise loops over the virtual blocks.
template <int DIM, bool ST, bool LIST>
__global__ void __launch_bounds__(kSweepThreads, LIST ? (ST ? MPHX_P1ST_MINB : MPHX_P1_MINB) : 1)
k_pass1_v3(const Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, int batch,
           double *__restrict__ P, double *__restrict__ volStrain, double *__restrict__ divP, double *__restrict__ densA,
           double *__restrict__ gcx, double *__restrict__ gcy, double *__restrict__ gcz, double *__restrict__ PA, PairList pl)
{
    const int n = ctl->n;
    if (n <= 0) return;
    const float filt2 = ctl->filt2;
    const int vblocks = (n + kSweepThreads - 1) / kSweepThreads;
    if constexpr (LIST) {
        if ((int)blockIdx.x >= vblocks) return;
        pass1_block<DIM, ST, true>(blockIdx.x, n, p, cellStart, g, ph, filt2, batch, P, volStrain, divP, densA, gcx, gcy, gcz, PA, pl);
    } else {
        if (pl.count && *reinterpret_cast<volatile const int *>(pl.flags) == 0) return; // no list overflowed in this step
        for (int vb = blockIdx.x; vb < vblocks; vb += gridDim.x) {
            pass1_block<DIM, ST, false>(vb, n, p, cellStart, g, ph, filt2, batch, P, volStrain, divP, densA, gcx, gcy, gcz, PA, pl);
            __syncthreads(); // the shared-memory queue is reused by the next virtual block
        }
    }
}

// K6 "pass 2": force sums (PressureP :2394-2425, PressureA :2225-2259, DiffuseInterface :2265-2312,
// ViscosityV :2480-2522, InterfaceForce :2439-2473) + gravity :2917 + explicit integration
// (:2938-2956, :1892-1907).
//   LIST = true : neighbours come from pass 1's list; particles whose list overflowed are skipped.
//   LIST = false: stencil sweep; with pl.count set only the overflowed particles are processed
//                 (and the block exits at once if it has none), otherwise all particles.
// measurement helpers (not part of the step) ------------------------------------------------------------------
// dense FP64 FMA throughput of the device: 8 independent chains per thread (the roofline's second denominator)
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0 + 1e-12, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 1.2345e300) out[0] = r; // (never true: keeps the chains alive)
}
// candidates in the current lists and pairs within the largest kernel radius (the algorithmic FP64 work of a sweep:
// SURVEY 8(d) counts ~15 flop per candidate examined and ~45 per in-radius pair)
__global__ void k_count_pairs(const Ctl *ctl, Particles p, GridDesc g, PairList pl, double r2max, unsigned long long *out /* [2] */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cand = 0, in = 0;
    if (i < ctl->n && pl.count[i] <= pl.L) {
        const int cnt = pl.count[i];
        const Rec a = p.ra[i];
        const MinImage mi(g);
        for (int k = 0; k < cnt; ++k) {
            const int j = pl.nbr[(size_t)k * pl.cap + i];
            const Rec b = p.ra[j];
            double dx = b.a - a.a, dy = b.b - a.b, dz = b.c - a.c;
            mi.apply(dx, dy, dz);
            if (j != i && dx * dx + dy * dy + dz * dz <= r2max) ++in;
        }
        cand = (unsigned long long)cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cand += __shfl_xor_sync(0xffffffffu, cand, o); in += __shfl_xor_sync(0xffffffffu, in, o); }
    if ((threadIdx.x & 31) == 0 && cand) { atomicAdd(&out[0], cand); atomicAdd(&out[1], in); }
}

template <int DIM, bool ST, bool LIST>
__device__ __forceinline__ void
pass2_block(Ctl *ctl, int vblock, int n, Particles p, const int *__restrict__ cellStart, const GridDesc &g, const Phys &ph, float filt2, int batch,
           const double *__restrict__ P, const double *__restrict__ PA, const double *__restrict__ gcx,
           const double *__restrict__ gcy, const double *__restrict__ gcz, double *__restrict__ ox,
           double *__restrict__ oy, double *__restrict__ oz, double *__restrict__ ovx, double *__restrict__ ovy,
           double *__restrict__ ovz, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
           double *__restrict__ ax, double *__restrict__ ay, double *__restrict__ az, Solid sol, PairList pl, Subset sub)
{
    __shared__ double s_visc[kTypeCount][kTypeCount];
    // pair viscosity table with the constant factors of the viscous term folded in:
    // c_d mu_ij V * (-cdv)   (:2505-2512, dwij = -dwvdr)
    for (int e = threadIdx.x; e < kTypeCount * kTypeCount; e += blockDim.x)
        s_visc[e / kTypeCount][e % kTypeCount] = -ph.viscpair[e / kTypeCount][e % kTypeCount] * ph.cdv;
    const int t0 = vblock * blockDim.x + threadIdx.x;
    const int i0 = sub.slots ? (t0 < sub.count ? sub.slots[t0] : n) : t0; // (a listed subset, or particle t0)
    const int i = i0 < n ? i0 : n - 1;
    const int tflag = p.type[i], ti = real_type(tflag), keyi = p.key[i];
    const bool solid_i = is_structure_type(ti);
    int mycount = 0;
    bool mine = i0 < n && !(sub.skip_solids && solid_i);
    if (pl.count) {
        mycount = pl.count[i];
        mine = mine && (LIST ? mycount <= pl.L : mycount > pl.L);
        if (!LIST && !__syncthreads_or(mine ? 1 : 0)) return; // no overflowed particle in this block
    }
    __syncthreads();
    const double xi = p.x[i], yi = p.y[i], zi = p.z[i];
    const double vxi = p.vx[i], vyi = p.vy[i], vzi = p.vz[i];
    const bool active = particle_active(g, i0, n, tflag, keyi);
    const double Pi = P[i];
    const double rp2 = ph.rp2, irp = ph.irp, rv2 = ph.rv2, irv = ph.irv;
    const double rpv2 = rp2 > rv2 ? rp2 : rv2;
    const double cpv = ph.cdp * ph.vol; // dwp/dr prefactor times particle volume
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    const Rec *__restrict__ RA = p.ra;
    const Rec *__restrict__ RB = p.rb;
    const double *visc_row = s_visc[ti];
    double PAi = 0.0, gi0 = 0.0, gi1 = 0.0, gi2 = 0.0, ai = 0.0;
    if (ST) { PAi = PA[i]; gi0 = gcx[i]; gi1 = gcy[i]; gi2 = gcz[i]; ai = ph.cofa[ti] * ph.cofk * ph.cofk; }
    const double gscale = ph.vol / ph.l0;
    // pair terms; rb = (vy, vz, PressureP, type bits) of j
    auto pair = [&](int j, double dx, double dy, double dz, double r2, double vxj, const Rec rb) {
#if !MPHX_BRANCHFREE
        if (j == i) return;
#endif
        const bool other = j != i;
        const int tj = (int)__double_as_longlong(rb.d);
        if (solid_i) {
            if (other && r2 < rp2 && !is_structure_type(tj)) { // :2455, :2447
                const double rinv = rsqrt_nr(r2);
                const double cc = (Pi + rb.c) * (1.0 - r2 * rinv * irp) * rinv * cpv;
                F0 += cc * dx; F1 += cc * dy; F2 += cc * dz;
            }
            return;
        }
#if MPHX_BRANCHFREE
        { // straight-line form (see pass 1): masked terms are exact zeros
            const bool inP = other && r2 < rp2, inV = other && r2 < rv2; // :2410, :2496 (strict)
            const double r2s = (inP || inV) ? r2 : 1.0;
            const double rinv = rsqrt_nr(r2s);
            const double r = r2s * rinv;
            const double cP = (Pi + rb.c) * (1.0 - r * irp) * rinv * cpv;
            const double ux = vxj - vxi, uy = rb.a - vyi, uz = rb.b - vzi;
            const double ue = (ux * dx + uy * dy + uz * dz) * rinv;
            const double cV = visc_row[tj] * ue * (1.0 - r * irv) * (rinv * rinv);
            double cc = inP ? cP : 0.0;
            cc += inV ? cV : 0.0;
            F0 += cc * dx; F1 += cc * dy; F2 += cc * dz;
        }
#else
        if (r2 < rpv2) {
            const bool inP = r2 < rp2, inV = r2 < rv2; // :2410, :2496 (strict)
            const double rinv = rsqrt_nr(r2);
            const double r = r2 * rinv;
            double cc = 0.0;
            if (inP) cc = (Pi + rb.c) * (1.0 - r * irp) * rinv * cpv;
            if (inV) {
                const double ux = vxj - vxi, uy = rb.a - vyi, uz = rb.b - vzi;
                const double ue = (ux * dx + uy * dy + uz * dz) * rinv;
                cc += visc_row[tj] * ue * (1.0 - r * irv) * (rinv * rinv);
            }
            F0 += cc * dx; F1 += cc * dy; F2 += cc * dz;
        }
#endif
        if (ST && other && r2 < ph.ra2) { // :2243, :2285 (RadiusG == RadiusA)
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
    };
    if (LIST) {
        const Bucket3 b = split_key<DIM>(g, active ? keyi : 0);
        const bool warp_wraps = __any_sync(0xffffffffu, active && mine && stencil_wraps<DIM>(g, b));
        const MinImage mi(g);
        const int cnt = (active && mine) ? mycount : 0;
        const int *lp = pl.nbr + i;
        const size_t ls = (size_t)pl.cap;
        int k = 0;
        int j0 = 0, j1 = 0;
        if (cnt > 0) j0 = __ldcs(lp);
        if (cnt > 1) j1 = __ldcs(lp + ls);
        for (; k + 1 < cnt; k += 2) { // two pairs per trip; the next two list entries are fetched first
            const int a_j = j0, b_j = j1;
            if (k + 2 < cnt) j0 = __ldcs(lp + (size_t)(k + 2) * ls);
            if (k + 3 < cnt) j1 = __ldcs(lp + (size_t)(k + 3) * ls);
            const Rec a0 = ld_rec_nc(RA + a_j), a1 = ld_rec_nc(RA + b_j);
            const Rec c0 = ld_rec_nc(RB + a_j), c1 = ld_rec_nc(RB + b_j);
            double dx0 = a0.a - xi, dy0 = a0.b - yi, dz0 = a0.c - zi;
            double dx1 = a1.a - xi, dy1 = a1.b - yi, dz1 = a1.c - zi;
            if (warp_wraps) { mi.apply(dx0, dy0, dz0); mi.apply(dx1, dy1, dz1); }
            pair(a_j, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0, a0.d, c0);
            pair(b_j, dx1, dy1, dz1, dx1 * dx1 + dy1 * dy1 + dz1 * dz1, a1.d, c1);
        }
        if (k < cnt) {
            const Rec a0 = ld_rec_nc(RA + j0), c0 = ld_rec_nc(RB + j0);
            double dx0 = a0.a - xi, dy0 = a0.b - yi, dz0 = a0.c - zi;
            if (warp_wraps) mi.apply(dx0, dy0, dz0);
            pair(j0, dx0, dy0, dz0, dx0 * dx0 + dy0 * dy0 + dz0 * dz0, a0.d, c0);
        }
    } else {
        __shared__ SweepShared sm;
        load_stencil(sm, g);
        __syncthreads();
        const bool go = active && mine;
        sweep<DIM>(sm, g, cellStart, p.pf, p.ra, i, go, go ? keyi : 0, xi, yi, zi, filt2, batch,
            [&](int j, double dx, double dy, double dz, double r2, double vxj) {
                pair(j, dx, dy, dz, r2, vxj, ld_rec_nc(RB + j));
            });
    }
    if (!mine) return;
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
            // (slab mode: the owner's values; the velocity is published to every rank by k_solid_publish_V, the force
            // stays with the owner, who reports it)
            const int s = p.id[i] - sol.sb;
            sol.vx[s] = nvx; sol.vy[s] = nvy; sol.vz[s] = nvz;
            sol.fx[s] = F0; sol.fy[s] = F1; sol.fz[s] = F2;
        }
    }
    if (!(nx - nx == 0.0) || !(ny - ny == 0.0) || !(nz - nz == 0.0)) atomicOr(&ctl->err, kErrNaN); // health flag: non-finite position
    ox[i] = nx; oy[i] = ny; oz[i] = nz; ovx[i] = nvx; ovy[i] = nvy; ovz[i] = nvz;
    fx[i] = F0; fy[i] = F1; fz[i] = F2; ax[i] = a0; ay[i] = a1; az[i] = a2;
}

template <int DIM, bool ST, bool LIST>
__global__ void __launch_bounds__(kSweepThreads, LIST ? (ST ? MPHX_P2ST_MINB : MPHX_P2_MINB) : 1)
k_pass2_v3(Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, int batch,
           const double *__restrict__ P, const double *__restrict__ PA, const double *__restrict__ gcx,
           const double *__restrict__ gcy, const double *__restrict__ gcz, double *__restrict__ ox,
           double *__restrict__ oy, double *__restrict__ oz, double *__restrict__ ovx, double *__restrict__ ovy,
           double *__restrict__ ovz, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
           double *__restrict__ ax, double *__restrict__ ay, double *__restrict__ az, Solid sol,
           PairList pl, Subset sub)
{
    const int n = ctl->n;
    if (n <= 0) return;
    const float filt2 = ctl->filt2;
    const int vblocks = ((sub.slots ? sub.count : n) + kSweepThreads - 1) / kSweepThreads;
    if constexpr (LIST) {
        if ((int)blockIdx.x >= vblocks) return;
        pass2_block<DIM, ST, true>(ctl, blockIdx.x, n, p, cellStart, g, ph, filt2, batch, P, PA, gcx, gcy, gcz, ox, oy, oz, ovx, ovy, ovz, fx, fy,
                                   fz, ax, ay, az, sol, pl, sub);
    } else {
        if (pl.count && *reinterpret_cast<volatile const int *>(pl.flags) == 0) return; // no list overflowed in this step
        for (int vb = blockIdx.x; vb < vblocks; vb += gridDim.x) {
            pass2_block<DIM, ST, false>(ctl, vb, n, p, cellStart, g, ph, filt2, batch, P, PA, gcx, gcy, gcz, ox, oy, oz, ovx, ovy, ovz, fx, fy,
                                        fz, ax, ay, az, sol, pl, sub);
            __syncthreads();
        }
    }
}

} // namespace mphx
