// sweep_pair.cuh -- k_filter2: the candidate filter (K5a) with TWO particles per thread.
//
// The one-particle filter (k_filter, sweep.cuh) is bound by the bytes the L1 returns to registers:
// every lane loads its own copy of a 7-bucket window of every stencil column although adjacent lanes'
// windows overlap by six buckets.  Particles 2t and 2t+1 are neighbours in the cell-sorted order
// (normally adjacent buckets of one bucket column), so thread t filters BOTH against one shared
// window per stencil column: each 24-byte pair record that reaches the registers is tested against two
// particles with the packed f32x2 instructions (register tiling), which cuts the returned bytes per
// particle by ~40 %.  Each particle keeps its own candidate list (same layout, same ascending order
// as k_filter writes), so pass 1 / pass 2 are unchanged.
//
// (Tried and rejected on B200, see DESIGN.md: one SHARED list per pair and two particles per thread in
// the physics kernels -- register pressure and the longer lists cost more than the saved gathers.)
#pragma once
#include <type_traits>

#include "sweep.cuh"

namespace mphx {

#ifndef MPHX_FILTER2_MINB
#define MPHX_FILTER2_MINB 8
#endif

// K5a, two particles per thread; same outputs as k_filter (one list per particle).
// BIG: the list has 2^32 or more slots (e.g. 10^8 particles on one GPU): 64-bit offsets.
template <int DIM, bool BIG>
__global__ void __launch_bounds__(kSweepThreads, MPHX_FILTER2_MINB)
k_filter2(const Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, PairList pl)
{
    if (!ctl->rebuild) return; // the list of an earlier step is still a superset of every cut-off set
    const int n = ctl->n;
    const float filt2 = ctl->filt2;
    using off_t = typename std::conditional<BIG, unsigned long long, unsigned>::type;
    __shared__ int s_dlo[kMaxStencil], s_dhi[kMaxStencil], s_sdx[kMaxStencil], s_sdy[kMaxStencil], s_sh[kMaxStencil];
    for (int e = threadIdx.x; e < g.nsten; e += blockDim.x) {
        const int dx = g.sdx[e], dy = g.sdy[e], h = g.sh[e];
        const int d = (DIM == 3) ? (dx * g.ny + dy) * g.nz : dx * g.ny;
        s_dlo[e] = d - h; s_dhi[e] = d + h + 1; s_sdx[e] = dx; s_sdy[e] = dy; s_sh[e] = h;
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = 2 * t, ib = ia + 1;
    if (ia >= n) return;
    const int keyA = p.key[ia], keyB = ib < n ? p.key[ib] : g.ncells;
    const bool actA = particle_active(g, ia, n, p.type[ia], keyA);
    const bool actB = ib < n && particle_active(g, ib, n, p.type[ib], keyB);
    if (!actA && !actB) { pl.count[ia] = 0; if (ib < n) pl.count[ib] = 0; return; }
    const PfPair *__restrict__ pf = p.pf;
    const PfPair own = ld_pf_nc(pf + t); // (x.x, y.x, z.x) = particle 2t, (x.y, y.y, z.y) = particle 2t+1
    const Bucket3 bA = split_key<DIM>(g, actA ? keyA : keyB), bB = split_key<DIM>(g, actB ? keyB : keyA);
    const bool wrapA = stencil_wraps<DIM>(g, bA), wrapB = stencil_wraps<DIM>(g, bB);

    int *__restrict__ nbr = pl.nbr;
    off_t stride = (off_t)pl.cap;
    off_t parkA = (off_t)pl.L * stride + (off_t)ia; // (particle 2t+1: parkA + 1)
    float filt = filt2;
    // keep the loop invariants in registers
    if constexpr (BIG) asm volatile("" : "+l"(nbr), "+l"(stride), "+l"(parkA), "+f"(filt));
    else asm volatile("" : "+l"(nbr), "+r"(stride), "+r"(parkA), "+f"(filt));
    off_t offA = (off_t)ia, offB = (off_t)ia + 1u;

    // one contiguous run [jb, je) of candidates against particle A (filter radius^2 fA; negative = masked)
    // and particle B (fB)
    auto scan_run = [&](int jb, int je, float ax, float ay, float az, float fA, float bx, float by, float bz, float fB) {
        const int len = je - jb, lenm1 = len - 1;
        const float2 nax = make_float2(-ax, -ax), nay = make_float2(-ay, -ay), naz = make_float2(-az, -az);
        const float2 nbx = make_float2(-bx, -bx), nby = make_float2(-by, -by), nbz = make_float2(-bz, -bz);
        int j0 = jb & ~1;
        int tt = j0 - jb; // -1 or 0: position of the pair's first element in the run
        const PfPair *pp = pf + (j0 >> 1);
        for (; tt < len; tt += 2, j0 += 2, ++pp) {
            const PfPair f = ld_pf_nc(pp);
            const float2 ax2 = __fadd2_rn(f.x, nax), ay2 = __fadd2_rn(f.y, nay), az2 = __fadd2_rn(f.z, naz);
            const float2 bx2 = __fadd2_rn(f.x, nbx), by2 = __fadd2_rn(f.y, nby), bz2 = __fadd2_rn(f.z, nbz);
            float2 da = __fmul2_rn(ax2, ax2), db = __fmul2_rn(bx2, bx2);
            da = __ffma2_rn(ay2, ay2, da); db = __ffma2_rn(by2, by2, db);
            da = __ffma2_rn(az2, az2, da); db = __ffma2_rn(bz2, bz2, db);
            const bool in0 = (unsigned)tt < (unsigned)len, in1 = tt < lenm1;
            if (in0 && da.x <= fA) { nbr[offA] = j0; offA = min(offA + stride, parkA); }
            if (in0 && db.x <= fB) { nbr[offB] = j0; offB = min(offB + stride, parkA + (off_t)1); }
            if (in1 && da.y <= fA) { nbr[offA] = j0 + 1; offA = min(offA + stride, parkA); }
            if (in1 && db.y <= fB) { nbr[offB] = j0 + 1; offB = min(offB + stride, parkA + (off_t)1); }
        }
    };
    // the stencil of ONE particle (general path: periodic images handled per column segment)
    auto scan_single = [&](int key, const Bucket3 &b, bool wraps, float fx, float fy, float fz, bool isA) {
        const float fA = isA ? filt : -1.f, fB = isA ? -1.f : filt;
        const int nsten = g.nsten;
        if (!wraps) {
            for (int e = 0; e < nsten; ++e) {
                const int jb = __ldg(cellStart + (key + s_dlo[e])), je = __ldg(cellStart + (key + s_dhi[e]));
                scan_run(jb, je, fx, fy, fz, fA, fx, fy, fz, fB);
            }
            return;
        }
        const int cx = b.cx, cy = b.cy, cr = b.cr, nr = b.nr;
        for (int e = 0; e < nsten; ++e) {
            int ccx = cx + s_sdx[e];
            float x = fx, y = fy, z = fz;
            if (ccx < 0) { ccx += g.nx; x += (float)g.nx; }
            else if (ccx >= g.nx) { ccx -= g.nx; x -= (float)g.nx; }
            int base;
            if (DIM == 3) {
                int ccy = cy + s_sdy[e];
                if (ccy < 0) { ccy += g.ny; y += (float)g.ny; }
                else if (ccy >= g.ny) { ccy -= g.ny; y -= (float)g.ny; }
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
                const float yy = (DIM == 2) ? y + shf : y;
                const float zz = (DIM == 3) ? z + shf : z;
                scan_run(cellStart[base + a], cellStart[base + bb + 1], x, yy, zz, fA, x, yy, zz, fB);
            }
        }
    };

    // joint scan: both particles active, stencils inside the box, same bucket column, buckets close along
    // the run axis -> per stencil column ONE window [first bucket of A's run, last bucket of B's run]
    const bool joint = actA && actB && !wrapA && !wrapB && bA.cx == bB.cx && bA.cy == bB.cy && (bB.cr - bA.cr) <= 4;
    if (joint) {
        const int nsten = g.nsten;
        int jbn = __ldg(cellStart + (keyA + s_dlo[0])), jen = __ldg(cellStart + (keyB + s_dhi[0]));
        for (int e = 0; e < nsten; ++e) {
            const int jb = jbn, je = jen;
            const int en = (e + 1 < nsten) ? e + 1 : e; // (the last trip re-reads its own bounds: no branch)
            jbn = __ldg(cellStart + (keyA + s_dlo[en])); jen = __ldg(cellStart + (keyB + s_dhi[en]));
            scan_run(jb, je, own.x.x, own.y.x, own.z.x, filt, own.x.y, own.y.y, own.z.y, filt);
        }
    } else {
        // separate scans (end of a bucket column, stencil across the periodic box, an inactive partner)
        if (actA) scan_single(keyA, bA, wrapA, own.x.x, own.y.x, own.z.x, true);
        if (actB) scan_single(keyB, bB, wrapB, own.x.y, own.y.y, own.z.y, false);
    }
    // (a list of exactly L entries counts as overflowed)
    if (!actA) pl.count[ia] = 0;
    else if (offA == parkA) { pl.count[ia] = pl.L + 1; atomicOr(pl.flags, 1); }
    else pl.count[ia] = (int)((offA - (off_t)ia) / stride);
    if (ib < n) {
        if (!actB) pl.count[ib] = 0;
        else if (offB == parkA + (off_t)1) { pl.count[ib] = pl.L + 1; atomicOr(pl.flags, 1); }
        else pl.count[ib] = (int)((offB - (off_t)ib) / stride);
    }
}

} // namespace mphx
