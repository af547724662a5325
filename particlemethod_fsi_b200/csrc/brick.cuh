// brick.cuh -- pass 1 with the neighbourhood staged in shared memory (the structure BASELINE.json's north star names:
// "TMA or shared-memory staging of each cell's 27-neighbour stencil", one block per BRICK of buckets).
//
// A brick is 8 x 8 x 8 buckets.  Its particles and those of the surrounding halo of `range` (3) buckets are, in the
// cell-sorted order, (8+6) x (8+6) contiguous z-runs of 8+6 buckets.  One block stages them in shared memory --
// the 32-byte (x, y, z, vx) records run by run with 1-D bulk async copies (cp.async.bulk + mbarrier, i.e. the TMA
// unit; the runs ARE contiguous byte ranges), (vy, vz) with plain loads -- and then every thread walks the candidate
// list of one owned particle reading its neighbours from shared memory: no L1 tag lookups, no L2 round trips inside
// the pair loop.  The candidate list is the step's list re-indexed to brick-local positions (k_brick_localize, rebuild
// steps only).  Bricks that touch the periodic box faces, or whose halo holds more particles than the staging buffer,
// stay with the global-gather list kernel (k_pass1_v3), as do particles whose list overflowed.
//
// Same pair arithmetic, same candidate order as k_pass1_v3: the two paths agree bit for bit.
#pragma once
#include "sweep.cuh"

namespace mphx {

constexpr int kBrick = 8;                 // buckets per brick edge
constexpr int kBrickMaxRange = 3;         // halo buckets (the stencil range) this layout is sized for
constexpr int kBrickRunsMax = (kBrick + 2 * kBrickMaxRange) * (kBrick + 2 * kBrickMaxRange); // 196 z-runs
constexpr int kBrickOwnRuns = kBrick * kBrick;                                              // 64 owned z-runs
constexpr int kBrickCap = 4096;           // staged particles (48 B each: 192 KB of the 227 KB a block may have)
constexpr int kBrickThreads = 512;

struct BrickGrid { int nbx, nby, nbz, nbricks; };
__host__ __device__ inline BrickGrid brick_grid(const GridDesc &g)
{
    BrickGrid b;
    b.nbx = (g.nx + kBrick - 1) / kBrick; b.nby = (g.ny + kBrick - 1) / kBrick; b.nbz = (g.nz + kBrick - 1) / kBrick;
    b.nbricks = b.nbx * b.nby * b.nbz;
    return b;
}

// The z-runs of one brick, computed by its block from the bucket offsets (valid until the next rebuild).
struct BrickRuns {
    int start[kBrickRunsMax];   // first sorted slot of halo run r = rx * nry + ry
    int off[kBrickRunsMax + 1]; // its first staged (brick-local) index; off[nruns] = staged particles
    int own_beg[kBrickOwnRuns], own_off[kBrickOwnRuns + 1]; // owned part of run (ox, oy): first slot, first owned index
    int bx0, by0, bz0, nry, nruns, edge;
};
__device__ __forceinline__ void brick_runs(BrickRuns &br, const GridDesc &g, const int *__restrict__ cellStart, int b)
{
    const BrickGrid bg = brick_grid(g);
    const int R = g.range;
    if (threadIdx.x == 0) {
        const int bbz = b % bg.nbz, t = b / bg.nbz;
        br.bz0 = bbz * kBrick; br.by0 = (t % bg.nby) * kBrick; br.bx0 = (t / bg.nby) * kBrick;
        br.nry = kBrick + 2 * R; br.nruns = br.nry * br.nry;
        // a halo that leaves the (local) bucket grid would need periodic images: those bricks keep the global path
        br.edge = (R > kBrickMaxRange) || br.bx0 - R < 0 || br.bx0 + kBrick + R > g.nx || br.by0 - R < 0 || br.by0 + kBrick + R > g.ny ||
                  br.bz0 - R < 0 || br.bz0 + kBrick + R > g.nz;
    }
    __syncthreads();
    if (br.edge) return;
    const int nry = br.nry, nruns = br.nruns;
    // run lengths and their exclusive scans (block-wide shuffles scans: blockDim >= 256 covers the 196 runs in one go)
    int len = 0, olen = 0;
    const int r = threadIdx.x;
    if (r < nruns) {
        const int cx = br.bx0 - R + r / nry, cy = br.by0 - R + r % nry;
        const int k0 = (cx * g.ny + cy) * g.nz + (br.bz0 - R);
        const int s0 = cellStart[k0];
        br.start[r] = s0;
        len = cellStart[k0 + kBrick + 2 * R] - s0;
    }
    if (r < kBrickOwnRuns) {
        const int cx = br.bx0 + r / kBrick, cy = br.by0 + r % kBrick;
        const int k0 = (cx * g.ny + cy) * g.nz + br.bz0;
        const int s0 = cellStart[k0];
        br.own_beg[r] = s0;
        olen = cellStart[k0 + kBrick] - s0;
    }
    __shared__ int tot, otot;
    const int ex = block_exclusive_scan(len, &tot);
    const int oex = block_exclusive_scan(olen, &otot);
    if (r < nruns) br.off[r] = ex;
    if (r < kBrickOwnRuns) br.own_off[r] = oex;
    if (r == 0) {
        br.off[nruns] = tot;
        br.own_off[kBrickOwnRuns] = otot;
        if (tot > kBrickCap) br.edge = 1; // too many particles to stage
    }
    __syncthreads();
}
// owned particle `t` of the brick -> (sorted slot, staged index)
__device__ __forceinline__ void brick_owned(const BrickRuns &br, int range, int t, int &slot, int &local)
{
    int lo = 0, hi = kBrickOwnRuns; // largest o with own_off[o] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (br.own_off[mid] <= t) lo = mid; else hi = mid;
    }
    slot = br.own_beg[lo] + (t - br.own_off[lo]);
    const int r = (lo / kBrick + range) * br.nry + (lo % kBrick + range);
    local = br.off[r] + (slot - br.start[r]);
}

// Rebuild steps: owned particles per brick (its exclusive scan, brick_base, is where the brick's rows of the staged
// list start: the list is stored in BRICK order so that the threads of a brick read consecutive entries)
__global__ void k_brick_count(const Ctl *ctl, const int *__restrict__ cellStart, GridDesc g, int nbricks, int *__restrict__ nown)
{
    if (!ctl->rebuild) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbricks) return;
    const BrickGrid bg = brick_grid(g);
    const int bz0 = (b % bg.nbz) * kBrick, t = b / bg.nbz, by0 = (t % bg.nby) * kBrick, bx0 = (t / bg.nby) * kBrick;
    const int z1 = min(bz0 + kBrick, g.nz);
    int cnt = 0;
    for (int cx = bx0; cx < min(bx0 + kBrick, g.nx); ++cx)
        for (int cy = by0; cy < min(by0 + kBrick, g.ny); ++cy) {
            const int k0 = (cx * g.ny + cy) * g.nz;
            cnt += cellStart[k0 + z1] - cellStart[k0 + bz0];
        }
    nown[b] = cnt;
}

// Rebuild steps: decide which bricks take the staged path (flag 1) and re-index their particles' candidate lists to
// staged positions (16 bit), stored at the particle's rank in brick order.  in_brick[i] = 1: particle i is handled by
// k_brick_pass1 (the list kernel skips it).
__global__ void __launch_bounds__(256)
k_brick_localize(const Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, PairList pl, unsigned short *__restrict__ lnbr,
                 const int *__restrict__ brick_base, unsigned char *__restrict__ brick_ok, unsigned char *__restrict__ in_brick)
{
    if (!ctl->rebuild) return;
    __shared__ BrickRuns br;
    const int b = blockIdx.x;
    brick_runs(br, g, cellStart, b);
    const int R = g.range;
    int nown = 0;
    if (br.edge) { // the owned particles of an edge brick: only the mask has to be written
        for (int o = threadIdx.x; o < kBrickOwnRuns; o += blockDim.x) {
            const int cx = br.bx0 + o / kBrick, cy = br.by0 + o % kBrick;
            if (cx >= g.nx || cy >= g.ny) continue;
            const int z1 = min(br.bz0 + kBrick, g.nz);
            const int k0 = (cx * g.ny + cy) * g.nz;
            for (int i = cellStart[k0 + br.bz0]; i < cellStart[k0 + z1]; ++i) in_brick[i] = 0;
        }
        if (threadIdx.x == 0) brick_ok[b] = 0;
        return;
    }
    nown = br.own_off[kBrickOwnRuns];
    if (threadIdx.x == 0) brick_ok[b] = nown > 0 ? 1 : 0; // (an empty brick launches a block that returns at once)
    const int per = g.ny * g.nz;
    const size_t base = (size_t)brick_base[b];
    for (int t = threadIdx.x; t < nown; t += blockDim.x) {
        int i, li;
        brick_owned(br, R, t, i, li);
        const int cnt = pl.count[i];
        const bool listed = cnt <= pl.L; // (an overflowed list belongs to the bucket-walking fall-back)
        in_brick[i] = listed ? 1 : 0;
        if (!listed) continue;
        for (int k = 0; k < cnt; ++k) {
            const int j = pl.nbr[(size_t)k * pl.cap + i];
            const int kj = p.key[j];
            const int cx = kj / per, cy = (kj / g.nz) % g.ny;
            const int r = (cx - (br.bx0 - R)) * br.nry + (cy - (br.by0 - R));
            lnbr[(size_t)k * pl.cap + base + t] = (unsigned short)(br.off[r] + (j - br.start[r]));
        }
    }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// K5 "pass 1", staged: VolStrainP, DivergenceP -> PressureP for the owned particles of one interior brick.
template <int DIM>
__global__ void __launch_bounds__(kBrickThreads, 1)
k_brick_pass1(const Ctl *ctl, Particles p, const int *__restrict__ cellStart, GridDesc g, Phys ph, PairList pl, const unsigned short *__restrict__ lnbr,
              const int *__restrict__ brick_base, const unsigned char *__restrict__ brick_ok, double *__restrict__ P, double *__restrict__ volStrain,
              double *__restrict__ divP)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Rec *sra = reinterpret_cast<Rec *>(smem_raw);                                   // [kBrickCap] x y z vx
    double2 *svb = reinterpret_cast<double2 *>(smem_raw + sizeof(Rec) * kBrickCap); // [kBrickCap] vy vz
    __shared__ BrickRuns br;
    __shared__ __align__(8) unsigned long long mbar;
    const int b = blockIdx.x;
    if (!brick_ok[b]) return;
    brick_runs(br, g, cellStart, b);
    if (br.edge) return; // (cannot happen for a brick marked ok; keeps the block uniform)
    const int nruns = br.nruns, total = br.off[nruns];
    if (br.own_off[kBrickOwnRuns] == 0) return;
    // ---- stage the halo: (x, y, z, vx) records by bulk async copies, one per run; (vy, vz) by plain loads ----------
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"((unsigned)(total * (int)sizeof(Rec))) : "memory");
    __syncthreads();
    for (int r = threadIdx.x; r < nruns; r += blockDim.x) {
        const int cnt = br.off[r + 1] - br.off[r];
        if (cnt > 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sra + br.off[r])),
                         "l"(p.ra + br.start[r]), "r"((unsigned)(cnt * (int)sizeof(Rec))), "r"(smem_u32(&mbar))
                         : "memory");
    }
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int r = warp; r < nruns; r += nwarps) {
            const int cnt = br.off[r + 1] - br.off[r], o = br.off[r], s = br.start[r];
            for (int e = lane; e < cnt; e += 32) {
                const double2 v = *reinterpret_cast<const double2 *>(&p.rb[s + e]); // (vy, vz): the first half of the record
                svb[o + e] = v;
            }
        }
    }
    { // wait for the bulk copies (phase 0 of the barrier), then for everybody's plain stores
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    __syncthreads();
    // ---- the pair loop: candidates from shared memory --------------------------------------------------------
    const int nown = br.own_off[kBrickOwnRuns];
    const double rp2 = ph.rp2, irp = ph.irp;
    for (int t = threadIdx.x; t < nown; t += blockDim.x) {
        int i, li;
        brick_owned(br, g.range, t, i, li);
        const int tflag = p.type[i], ti = real_type(tflag);
        const int cnt = pl.count[i];
        if (cnt > pl.L || !particle_active(g, i, ctl->n, tflag, p.key[i])) continue;
        const Rec own = sra[li];
        const double2 ownv = svb[li];
        const double xi = own.a, yi = own.b, zi = own.c, vxi = own.d, vyi = ownv.x, vzi = ownv.y;
        double nP = 0.0, dv = 0.0;
        const unsigned short *lp = lnbr + (size_t)brick_base[b] + t; // (brick order: the threads of a warp read consecutive entries)
        const size_t ls = (size_t)pl.cap;
        auto pair = [&](int lj) { // same straight-line form as k_pass1_v3 (a masked pair contributes exact zeros)
            const Rec a = sra[lj];
            const double2 bv = svb[lj];
            const double dx = a.a - xi, dy = a.b - yi, dz = a.c - zi;
            const double r2 = dx * dx + dy * dy + dz * dz;
            const bool in = lj != li && r2 <= rp2; // :2333, :2362
            const double r2s = in ? r2 : 1.0;
            const double rinv = rsqrt_nr(r2s);
            const double q = in ? 1.0 - (r2s * rinv) * irp : 0.0;
            nP += q * q;
            const double ux = a.d - vxi, uy = bv.x - vyi, uz = bv.y - vzi;
            dv -= (ux * dx + uy * dy + uz * dz) * rinv * q;
        };
        // the list entries of the NEXT four pairs are in flight while four pairs are evaluated (the list streams from HBM;
        // the neighbours themselves come from shared memory)
        int jn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) jn[u] = u < cnt ? (int)__ldcs(lp + (size_t)u * ls) : li;
        for (int k = 0; k < cnt; k += 4) {
            int jc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) jc[u] = jn[u];
#pragma unroll
            for (int u = 0; u < 4; ++u) jn[u] = k + 4 + u < cnt ? (int)__ldcs(lp + (size_t)(k + 4 + u) * ls) : li;
#pragma unroll
            for (int u = 0; u < 4; ++u) pair(jc[u]); // (a padding entry is the particle itself: masked, exact zeros)
        }
        nP *= ph.cwp;
        dv *= ph.cdp;
        const double vs = nP - ph.n0p;                        // :2339
        const double kappa = (vs < 0.0) ? 0.0 : ph.bulk[ti]; // :2112-2113
        double pr = -ph.lambda[ti] * dv;                      // :2388
        if (vs > 0.0) pr += kappa * vs;                       // :2389-2391
        P[i] = pr; volStrain[i] = vs; divP[i] = dv;
        p.rb[i].c = pr; // the gather record pass 2 reads
    }
}

} // namespace mphx
