// mphx.cu -- context, step orchestration and the extern-"C" layer (include/mphx.h).
//
// One context drives one B200.  All state lives on the device as cell-sorted SoA; the host only
// keeps scalars (Time, wall centres) and enqueues kernels on the context's stream.
// There is no CPU fallback: every compute entry point fails with MPHX_ERR_NO_DEVICE / MPHX_ERR_CUDA
// when no sm_100 device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "sweep.cuh"
#include "sweep_pair.cuh"
#include "mphx.h"
#include "mphx_internal.h"

namespace mphx {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
            return MPHX_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

constexpr int kBlock = 128;
static inline int nblk(long long n, int b = kBlock) { return (int)((n + b - 1) / b); }

struct Ctx {
    mphx_params p;
    mphx_constants c;
    int device = 0;
    cudaStream_t stream = nullptr;
    int n = 0;        // particle slots currently held (slab mode: owned + ghosts + all solids; changes every step)
    int n_global = 0; // particles of the whole case (= n without slabs)
    int cap = 0;      // allocated slots
    int nf = 0, ns = 0, nw = 0; // class counts of the whole case
    int ranges[6];
    // slab mode (multi-GPU): see slab.inc
    bool slab = false;
    int rank = 0, nranks = 1, col_lo = 0, col_hi = 0, msg_cap = 0;
    int ghost_base[2] = {0, 0}, ghost_cnt[2] = {0, 0}, halo_cnt[2] = {0, 0};
    int *where = nullptr, *haloSrc[2] = {nullptr, nullptr}, *d_err = nullptr;
    int slab_err[4] = {0, 0, 0, 0}; // host copy of d_err, refreshed by stage_sort
    bool external_stream = false;
    bool uploaded = false, inited = false, surface_tension = false;
    double time = 0.0;
    double wall_center[kTypeCount][3];
    long long launches = 0;
    long long steps_done = 0;

    GridDesc grid;
    Phys phys;
    Particles S{}, T{}; // S: current cell-sorted state, T: scratch (permute target / pre-step positions)
    double *bx = nullptr, *by = nullptr, *bz = nullptr; // positions the buckets were built from
    int *cellCount = nullptr, *cellStart = nullptr, *slot = nullptr, *tmpIdx = nullptr, *blockSums = nullptr;
    int scan_blocks = 0;
    double *P = nullptr, *volStrain = nullptr, *divP = nullptr;
    double *densA = nullptr, *gcx = nullptr, *gcy = nullptr, *gcz = nullptr, *PA = nullptr;
    double *fx = nullptr, *fy = nullptr, *fz = nullptr, *ax = nullptr, *ay = nullptr, *az = nullptr;
    Solid sol{};
    double *d_inv_density = nullptr;
    double cw_tl = 0.0; // weight() prefactor (1.0/Swp)*(1.0/RP^d), src/main.cpp:291/293
    double *stage3a = nullptr, *stage3b = nullptr, *stage1 = nullptr, *stage9 = nullptr; // AoS staging (lazy)
    int *stagei = nullptr, *mask = nullptr, *mask2 = nullptr, *tmpi = nullptr;
    // compact owned-particle I/O (mphx_download_owned / mphx_upload_owned)
    int *own_scan = nullptr, *own_sums = nullptr, *own_slot = nullptr, *own_ids = nullptr;
    double *own_x = nullptr, *own_v = nullptr;
    int own_blocks = 0, own_count = -1;
    long long own_epoch = -1; // steps_done at the last mphx_download_owned
    bool buckets_valid = false;
    int sweep_batch = 12;  // stencil columns per filter/drain batch (3D)
    PairList pl{};         // pass 1 -> pass 2 neighbour list (nbr == nullptr: disabled, pass 2 sweeps again)
    int list_cap = -1;     // list slots per particle (-1: default by dimension, 0: no list)
    bool filter2 = true;   // build the lists two particles per thread (k_filter2, sweep_pair.cuh)
    // the solid sub-steps only need the solids' share of pass 2: they run on a second stream while the
    // fluid's share (FP64 / issue bound; the sub-steps are HBM bound) is still being computed
    bool overlap_solid = true;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_solid_ready = nullptr, ev_solid_done = nullptr;
    bool solids_pending = false;  // sub-steps enqueued on `side`, not yet joined by the main stream
    WallMotion held_wm{};         // arguments of the deferred solid share of the pre-step
    int held_wrap = 0;
    std::vector<void *> allocs;

    // phase timers (src/main.cpp:695-700 split)
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    double ms[5] = {0, 0, 0, 0, 0}; // rebuild, filter, pass 1, pass 2, solid sub-steps
    cudaEvent_t tev[2] = {nullptr, nullptr};

    template <class Tp> int alloc(Tp **ptr, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(Tp));
        if (e != cudaSuccess) {
            set_last_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
            return MPHX_ERR_NOMEM;
        }
        allocs.push_back(q);
        *ptr = (Tp *)q;
        return MPHX_OK;
    }
};

#define LAUNCH_ON(ctx, strm, kernel, grid, block, ...)                                             \
    do {                                                                                           \
        const int grid_ = (grid);                                                                  \
        if (grid_ > 0) {                                                                           \
            kernel<<<grid_, (block), 0, (strm)>>>(__VA_ARGS__);                                     \
            ++(ctx)->launches;                                                                     \
        }                                                                                          \
    } while (0)
#define LAUNCH(ctx, kernel, grid, block, ...)                                                      \
    do {                                                                                           \
        const int grid_ = (grid);                                                                  \
        if (grid_ > 0) { /* an empty slab launches nothing */                                      \
            kernel<<<grid_, (block), 0, (ctx)->stream>>>(__VA_ARGS__);                              \
            ++(ctx)->launches;                                                                     \
        }                                                                                          \
    } while (0)

static int alloc_particles(Ctx *c, Particles *p, size_t n)
{
    int rc = 0;
    rc |= c->alloc(&p->x, n); rc |= c->alloc(&p->y, n); rc |= c->alloc(&p->z, n);
    rc |= c->alloc(&p->vx, n); rc |= c->alloc(&p->vy, n); rc |= c->alloc(&p->vz, n);
    rc |= c->alloc(&p->type, n); rc |= c->alloc(&p->id, n); rc |= c->alloc(&p->key, n);
    rc |= c->alloc(&p->pf, n / 2 + 4);
    return rc ? MPHX_ERR_NOMEM : MPHX_OK;
}
// the gather records: one buffer serves both ping-pong sets (the permute reads SoA, writes records)
static int alloc_records(Ctx *c, Particles *a, Particles *b, size_t n)
{
    int rc = 0;
    rc |= c->alloc(&a->ra, n); rc |= c->alloc(&a->rb, n);
    b->ra = a->ra; b->rb = a->rb;
    return rc ? MPHX_ERR_NOMEM : MPHX_OK;
}

// stencil columns: all (dx,dy) within `range` whose buckets can hold a particle within the list
// cut-off of some particle of the home bucket; per column the half-length along the run axis.
static int build_stencil(GridDesc &g, double cutoff)
{
    const int R = g.range;
    const double cut2 = (cutoff / g.cellw) * (cutoff / g.cellw) * (1.0 + 1e-9) + 1e-9;
    auto gap = [](int d) { const int a = std::abs(d) - 1; return a > 0 ? (double)a : 0.0; };
    int ns = 0;
    if (g.dim == 3) {
        for (int dx = -R; dx <= R; ++dx)
            for (int dy = -R; dy <= R; ++dy) {
                const double m2 = gap(dx) * gap(dx) + gap(dy) * gap(dy);
                if (m2 > cut2) continue;
                int h = 0;
                for (int dz = 0; dz <= R; ++dz)
                    if (m2 + gap(dz) * gap(dz) <= cut2) h = dz;
                if (ns >= kMaxStencil) return MPHX_ERR_UNSUPPORTED;
                g.sdx[ns] = (signed char)dx; g.sdy[ns] = (signed char)dy; g.sh[ns] = (signed char)h;
                ++ns;
            }
    } else {
        for (int dx = -R; dx <= R; ++dx) {
            const double m2 = gap(dx) * gap(dx);
            if (m2 > cut2) continue;
            int h = 0;
            for (int dy = 0; dy <= R; ++dy)
                if (m2 + gap(dy) * gap(dy) <= cut2) h = dy;
            if (ns >= kMaxStencil) return MPHX_ERR_UNSUPPORTED;
            g.sdx[ns] = (signed char)dx; g.sdy[ns] = 0; g.sh[ns] = (signed char)h;
            ++ns;
        }
    }
    g.nsten = ns;
    return MPHX_OK;
}

static int setup_constants(Ctx *c)
{
    const mphx_params &p = c->p;
    int rc = mphx_compute_constants(&p, &c->c);
    if (rc) return rc;
    const mphx_constants &k = c->c;
    GridDesc &g = c->grid;
    std::memset(&g, 0, sizeof(g));
    g.dim = p.dim;
    g.nx = k.cell_count[0]; g.ny = k.cell_count[1]; g.nz = k.cell_count[2];
    g.ncells = k.cell_counts;
    g.range = k.stencil_range;
    g.cellw = k.cell_width;
    for (int d = 0; d < 3; ++d) { g.mn[d] = p.domain_min[d]; g.W[d] = k.domain_width[d]; }
    g.slab = 0; g.nxg = g.nx; g.xoff = 0; g.mn0g = p.domain_min[0];
    const double cutoff = k.max_radius + 0.1 * p.particle_spacing; // MaxRadius+MARGIN (:116,:1765)
    rc = build_stencil(g, cutoff);
    if (rc) { set_last_error("stencil too large (radius ratio too big)"); return rc; }
    // every traversed axis must hold the whole stencil once (otherwise the reference itself visits
    // buckets several times, :1751-1755)
    const int need = 2 * g.range + 1;
    if (g.nx < need || g.ny < need || (p.dim == 3 && g.nz < need)) {
        set_last_error("domain narrower than the neighbour stencil");
        return MPHX_ERR_UNSUPPORTED;
    }
    Phys &ph = c->phys;
    std::memset(&ph, 0, sizeof(ph));
    const bool two_d = p.dim == 2;
    auto hd = [&](double h) { return two_d ? h * h : h * h * h; };
    ph.dt = p.dt; ph.vol = k.particle_volume; ph.l0 = p.particle_spacing;
    ph.rp2 = k.radius_p * k.radius_p; ph.irp = 1.0 / k.radius_p;
    ph.cwp = 1.0 / k.swp * 1.0 / hd(k.radius_p); ph.cdp = ph.cwp * (-2.0 / k.radius_p);
    ph.rv2 = k.radius_v * k.radius_v; ph.irv = 1.0 / k.radius_v;
    ph.cdv = (1.0 / k.swv * 1.0 / hd(k.radius_v)) * (-2.0 / k.radius_v);
    ph.ra2 = k.radius_a * k.radius_a; ph.ira = 1.0 / k.radius_a;
    ph.cwa = 1.0 / k.swa * 1.0 / hd(k.radius_a);
    ph.cwg = 1.0 / k.swg * 1.0 / hd(k.radius_g); ph.cdg = ph.cwg * (-2.0 / k.radius_g);
    ph.r2g = k.r2g; ph.rg = k.radius_g;
    ph.n0p = k.n0p; ph.n0a = k.n0a; ph.cofk = k.cof_k;
    const double cd = two_d ? 8.0 : 10.0; // :2510 / :2512
    c->surface_tension = false;
    for (int t = 0; t < kTypeCount; ++t) {
        ph.mass[t] = p.density[t] * k.particle_volume; // :2105
        ph.inv_density[t] = 1.0 / p.density[t];        // :2877
        ph.bulk[t] = p.bulk_modulus[t];
        ph.lambda[t] = p.bulk_viscosity[t];
        ph.cofa[t] = k.cof_a[t];
        if (k.cof_a[t] != 0.0) c->surface_tension = true;
        for (int u = 0; u < kTypeCount; ++u) {
            const double mi = p.shear_viscosity[t], mj = p.shear_viscosity[u];
            ph.viscpair[t][u] = cd * (2.0 * (mi * mj) / (mi + mj)) * k.particle_volume; // :2505
            ph.ratio[t][u] = p.interaction_ratio[t][u];
        }
    }
    for (int d = 0; d < 3; ++d) ph.g[d] = p.gravity[d];
    c->cw_tl = (1.0 / k.swp) * (1.0 / hd(k.radius_p));
    return MPHX_OK;
}

// The solid sub-steps run on a second stream and are joined as late as possible: they overlap the fluid's
// share of pass 2 (single context) and the pre-step / migration / halo phases of the next step.
static bool defer_solids(const Ctx *c) { return c->overlap_solid && c->ns > 0 && c->side != nullptr && c->pl.nbr != nullptr; }
static int join_solids(Ctx *c)
{
    if (c->solids_pending) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_solid_done, 0));
        c->solids_pending = false;
    }
    return MPHX_OK;
}

// ---- bucket rebuild: K1 key/count, K2 scan, K3 scatter, K4 permute -------------------------------
static int stage_prestep(Ctx *c, bool prestep_motion, SlabSend snd)
{
    const int n = c->n;
    CK(cudaMemsetAsync(c->cellCount, 0, sizeof(int) * ((size_t)c->grid.ncells + 2), c->stream));
    WallMotion wm;
    std::memset(&wm, 0, sizeof(wm));
    wm.dt = c->p.dt;
    wm.active = (prestep_motion && c->time < 0.2 && c->nw > 0) ? 1 : 0; // Q7 (:3037)
    if (wm.active)
        for (int t = 0; t < kTypeCount; ++t)
            for (int d = 0; d < 3; ++d) {
                wm.center[t][d] = c->wall_center[t][d];
                wm.vel[t][d] = c->p.wall_velocity[t][d];
                wm.omega[t][d] = c->p.wall_omega[t][d];
                for (int e = 0; e < 3; ++e) wm.R[t][d][e] = c->c.wall_rotation[t][d][e];
            }
    if (n > 0)
        LAUNCH(c, k_prestep, nblk(n), kBlock, n, c->S, c->sol, c->grid, wm, prestep_motion ? 1 : 0, c->cellCount, c->slot, snd,
               (const int *)nullptr, defer_solids(c) ? 1 : 0);
    c->held_wm = wm; c->held_wrap = prestep_motion ? 1 : 0;
    if (prestep_motion) // :3066-3070 (host mirror of the wall centres)
        for (int t = 4; t < kTypeCount; ++t)
            for (int d = 0; d < 3; ++d) c->wall_center[t][d] += c->p.wall_velocity[t][d] * c->p.dt;
    CK(cudaGetLastError());
    return MPHX_OK;
}

// K2 scan, K3 scatter, K4 permute over the c->n slots keyed so far; afterwards c->n = live slots
static int stage_sort(Ctx *c)
{
    const int n = c->n;
    if (defer_solids(c)) { // the solids' share of the pre-step, once their sub-steps have finished
        int rc = join_solids(c);
        if (rc) return rc;
        SlabSend none{};
        LAUNCH(c, k_prestep, nblk(c->ns), kBlock, c->ns, c->S, c->sol, c->grid, c->held_wm, c->held_wrap, c->cellCount, c->slot, none,
               (const int *)c->sol.slot, 0);
    }
    const int nc = c->grid.ncells + 2; // + parked + dead buckets
    LAUNCH(c, k_scan_reduce, c->scan_blocks, kScanThreads, c->cellCount, nc, c->blockSums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, c->blockSums, c->scan_blocks);
    LAUNCH(c, k_scan_apply, c->scan_blocks, kScanThreads, c->cellCount, nc, c->blockSums, c->cellStart);
    if (n > 0) {
        LAUNCH(c, k_scatter_index, nblk(n), kBlock, n, c->S.key, c->slot, c->cellStart, c->tmpIdx);
        LAUNCH(c, k_permute, nblk(n), kBlock, n, c->S, c->T, c->cellStart, c->tmpIdx, c->grid, c->where, c->sol.slot, c->sol.sb);
    }
    std::swap(c->S, c->T);
    c->bx = c->S.x; c->by = c->S.y; c->bz = c->S.z;
    c->buckets_valid = true;
    if (c->slab) { // ghosts of the last step and emigrants sit in the dead bucket: drop them
        // (one synchronisation: the live slot count and the exchange error flags come back together)
        int keep = 0;
        CK(cudaMemcpyAsync(&keep, c->cellStart + c->grid.ncells + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->slab_err, c->d_err, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->n = keep;
    }
    CK(cudaGetLastError());
    return MPHX_OK;
}

static int rebuild_buckets(Ctx *c, bool prestep_motion)
{
    if (c->slab) { set_last_error("slab contexts are stepped through the mphx_slab_* calls"); return MPHX_ERR_INVALID; }
    SlabSend none{};
    int rc = stage_prestep(c, prestep_motion, none);
    if (rc) return rc;
    return stage_sort(c);
}

// squared cut-off (bucket units) of the sweep's fp32 filter: the exact cut-off plus a margin that
// covers the rounding of the fp32 bucket coordinates (<= 2 ulp at the largest coordinate, per axis)
// and of the fp32 distance arithmetic, so the filter is a superset of the fp64 predicate.
static float filter_radius2(const Ctx *c, double rmax)
{
    const GridDesc &g = c->grid;
    const double R = rmax / g.cellw;
    const double big = (double)std::max(std::max(g.nx, g.ny), std::max(g.nz, 4)) + 2.0 * g.range + 2.0;
    const double delta = std::ldexp(big, -22);           // 2 ulp_f32(big) per coordinate, both particles
    const double margin = 2.0 * std::sqrt(3.0) * (R + 1.0) * (2.0 * delta) + 3.0 * (2.0 * delta) * (2.0 * delta) +
                          8.0 * std::ldexp((R + 1.0) * (R + 1.0), -23) + 1e-6;
    return (float)(R * R + margin) * (1.0f + 1e-6f);
}

static float sweep_filter2(const Ctx *c)
{
    // the filter covers every kernel radius of BOTH passes: they share one candidate list
    const mphx_constants &k = c->c;
    double rmax = std::max(k.radius_p, k.radius_v);
    if (c->surface_tension) rmax = std::max(rmax, k.radius_a);
    return filter_radius2(c, rmax);
}

static void timer_mark(Ctx *c);
static int run_pass1(Ctx *c, bool timed = false)
{
    const int n = c->n;
    {
        const float f2 = sweep_filter2(c);
        const int batch = c->p.dim == 3 ? c->sweep_batch : c->grid.nsten;
        if (c->pl.nbr) { // K5a: candidate list of this step
            CK(cudaMemsetAsync(c->pl.flags, 0, sizeof(int), c->stream));
            const int npairs = (n + 1) / 2;
            // 32-bit offsets must hold a parked top plus one stride ((L + 1) * cap + i + cap) without wrapping
            const bool big = ((size_t)c->pl.L + 2) * (size_t)c->pl.cap >= 0xffffffffull;
            if (c->filter2 || big) {
#define F2(D, B) LAUNCH(c, (k_filter2<D, B>), nblk(npairs, kSweepThreads), kSweepThreads, n, c->S, c->cellStart, c->grid, f2, c->pl)
                if (c->p.dim == 3) { if (big) F2(3, true); else F2(3, false); }
                else               { if (big) F2(2, true); else F2(2, false); }
#undef F2
            } else {
                if (c->p.dim == 3) LAUNCH(c, k_filter<3>, nblk(n, kSweepThreads), kSweepThreads, n, c->S, c->cellStart, c->grid, f2, c->pl);
                else               LAUNCH(c, k_filter<2>, nblk(n, kSweepThreads), kSweepThreads, n, c->S, c->cellStart, c->grid, f2, c->pl);
            }
        }
        if (timed) timer_mark(c);
        // fused-sweep fall-backs: a small persistent grid when they only have to look at the overflow flag
        const int vblocks = nblk(n, kSweepThreads);
#define SWEEP_GRID(LIST) ((LIST) || !c->pl.nbr ? vblocks : std::min(vblocks, 4 * 148))
#define P1(D, ST, LIST) LAUNCH(c, (k_pass1_v3<D, ST, LIST>), SWEEP_GRID(LIST), kSweepThreads, vblocks, n, c->S, c->cellStart, c->grid, c->phys, f2, \
                         batch, c->P, c->volStrain, c->divP, c->densA, c->gcx, c->gcy, c->gcz, c->PA, c->pl)
#define P1D(ST, LIST) do { if (c->p.dim == 3) P1(3, ST, LIST); else P1(2, ST, LIST); } while (0)
        if (c->pl.nbr) { // list traversal, then the fused sweep for particles whose list overflowed (normally none)
            if (c->surface_tension) P1D(true, true); else P1D(false, true);
        }
        if (c->surface_tension) P1D(true, false); else P1D(false, false);
#undef P1D
#undef P1
    }
    CK(cudaGetLastError());
    return MPHX_OK;
}

// single context with solids and a candidate list: pass 2 is split so that the sub-steps can overlap it
static bool solid_split(const Ctx *c) { return defer_solids(c) && !c->slab; }

static int run_pass2(Ctx *c, double *solbuf = nullptr)
{
    const int n = c->n;
    {
        const float f2 = sweep_filter2(c);
        const int batch = c->p.dim == 3 ? c->sweep_batch : c->grid.nsten;
        const int vblocks = nblk(n, kSweepThreads);
#define P2(D, ST, LIST, GRID, SUB) LAUNCH(c, (k_pass2_v3<D, ST, LIST>), GRID, kSweepThreads, vblocks, n, c->S, c->cellStart, c->grid, \
                         c->phys, f2, batch, c->P, c->PA, c->gcx, c->gcy, c->gcz, c->T.x, c->T.y, c->T.z, c->T.vx, c->T.vy, c->T.vz, c->fx, \
                         c->fy, c->fz, c->ax, c->ay, c->az, c->sol, solbuf, c->pl, SUB)
#define P2D(ST, LIST, GRID, SUB) do { if (c->p.dim == 3) P2(3, ST, LIST, GRID, SUB); else P2(2, ST, LIST, GRID, SUB); } while (0)
#define P2ALL(GRIDL, SUB) do { \
        if (c->pl.nbr) { /* list traversal, then the sweep variant for particles whose list overflowed (normally none) */ \
            if (c->surface_tension) P2D(true, true, GRIDL, SUB); else P2D(false, true, GRIDL, SUB); \
        } \
        if (c->surface_tension) P2D(true, false, SWEEP_GRID(false), SUB); else P2D(false, false, SWEEP_GRID(false), SUB); \
    } while (0)
        if (solid_split(c)) {
            // the solids' share first (a few blocks), so that the sub-steps can start; then everything else
            const Subset solids{c->sol.slot, c->ns, 0}, rest{nullptr, 0, 1};
            P2ALL(nblk(c->ns, kSweepThreads), solids);
            CK(cudaEventRecord(c->ev_solid_ready, c->stream));
            P2ALL(vblocks, rest);
        } else {
            const Subset all{nullptr, 0, 0};
            P2ALL(vblocks, all);
        }
#undef P2ALL
#undef P2D
#undef P2
    }
    // S keeps type/id/key of this step's order and takes the integrated x,v; T keeps the pre-step
    // (bucket) positions for the neighbour-count diagnostics.
    std::swap(c->S.x, c->T.x); std::swap(c->S.y, c->T.y); std::swap(c->S.z, c->T.z);
    std::swap(c->S.vx, c->T.vx); std::swap(c->S.vy, c->T.vy); std::swap(c->S.vz, c->T.vz);
    c->bx = c->T.x; c->by = c->T.y; c->bz = c->T.z;
    CK(cudaGetLastError());
    return MPHX_OK;
}

static int run_solid_substeps(Ctx *c, cudaStream_t strm)
{
    if (c->ns <= 0) return MPHX_OK;
    const int substeps = (int)(c->p.dt / c->p.elastic_dt + 0.5); // :653
    const mphx_constants &k = c->c;
    const double cw = c->cw_tl;
    const int ns = c->ns;
    const int dbl = (c->p.ref_compat & MPHX_COMPAT_DOUBLE_UPDATE) ? 1 : 0;
    for (int s = 0; s < substeps; ++s) {
        if (c->p.dim == 3) {
            LAUNCH_ON(c, strm, k_solid_pass1<3>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, cw);
            LAUNCH_ON(c, strm, k_solid_pass2<3>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, cw,
                   c->p.elastic_dt, c->p.clamp_module, dbl, c->d_inv_density);
        } else {
            LAUNCH_ON(c, strm, k_solid_pass1<2>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, cw);
            LAUNCH_ON(c, strm, k_solid_pass2<2>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, cw,
                   c->p.elastic_dt, c->p.clamp_module, dbl, c->d_inv_density);
        }
    }
    CK(cudaGetLastError());
    return MPHX_OK;
}

static void timer_mark(Ctx *c)
{
    if (!c->timing) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->ev.push_back(e);
}
static void timer_resolve(Ctx *c)
{
    if (c->ev.empty()) return;
    cudaStreamSynchronize(c->stream);
    // events come in groups of 6 per step: start, after rebuild, after the filter, after pass 1, after
    // pass 2, after the solid sub-steps
    for (size_t i = 0; i + 5 < c->ev.size(); i += 6)
        for (int k = 0; k < 5; ++k) {
            float a = 0;
            cudaEventElapsedTime(&a, c->ev[i + k], c->ev[i + k + 1]);
            c->ms[k] += a;
        }
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    c->ev.clear();
}

static int one_step(Ctx *c, bool fluid_only)
{
    int rc;
    timer_mark(c);
    if ((rc = rebuild_buckets(c, true))) return rc; // calculateWall, PeriodicBoundary, resets, calculateNeighbor
    timer_mark(c);
    if ((rc = run_pass1(c, true))) return rc;       // filter; DensityA..DivergenceP, coefficients, PressureP/A
    timer_mark(c);
    if ((rc = run_pass2(c))) return rc;             // force sums, gravity, interface, acceleration, convection
    timer_mark(c);
    if (!fluid_only) {
        if (solid_split(c)) { // sub-steps on the second stream: joined by the next bucket sort (or by any reader)
            CK(cudaStreamWaitEvent(c->side, c->ev_solid_ready, 0));
            if ((rc = run_solid_substeps(c, c->side))) return rc;
            CK(cudaEventRecord(c->ev_solid_done, c->side));
            c->solids_pending = true;
        } else if ((rc = run_solid_substeps(c, c->stream))) return rc;
        c->time += c->p.dt; // :685
        ++c->steps_done;
    }
    timer_mark(c);
    if (c->ev.size() >= 6000) timer_resolve(c);
    return MPHX_OK;
}

// ---- initial structure lists (calculateInitialNeighbor :1497-1644) + Lame + Normalizer ----------
static int exact_lists(Ctx *c, bool structure_only, bool xy_only, int row_base, int nrows,
                       std::vector<long long> &offsets, int **d_ids_out, long long *total_out)
{
    // neighbour sets over the positions the buckets were last built from
    const int n = c->n;
    int *d_counts = nullptr;
    long long *d_off = nullptr;
    CK(cudaMalloc(&d_counts, sizeof(int) * (size_t)std::max(nrows, 1)));
    CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)std::max(nrows, 1), c->stream));
    const double cut = c->c.max_radius + 0.1 * c->p.particle_spacing;
    const double cutoff2 = cut * cut; // (MaxRadius+MARGIN)*(MaxRadius+MARGIN) :1765
    const Particles &S = c->S;
    if (c->p.dim == 3)
        LAUNCH(c, (k_neighbors_exact<3, 0>), nblk(n), kBlock, n, c->bx, c->by, c->bz, S.type, S.id, S.key, c->cellStart, c->grid,
               cutoff2, structure_only ? 1 : 0, xy_only ? 1 : 0, row_base, d_counts, (const long long *)nullptr, (int *)nullptr);
    else
        LAUNCH(c, (k_neighbors_exact<2, 0>), nblk(n), kBlock, n, c->bx, c->by, c->bz, S.type, S.id, S.key, c->cellStart, c->grid,
               cutoff2, structure_only ? 1 : 0, xy_only ? 1 : 0, row_base, d_counts, (const long long *)nullptr, (int *)nullptr);
    std::vector<int> counts((size_t)std::max(nrows, 1));
    CK(cudaMemcpyAsync(counts.data(), d_counts, sizeof(int) * (size_t)std::max(nrows, 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    offsets.assign((size_t)nrows + 1, 0);
    for (int r = 0; r < nrows; ++r) offsets[r + 1] = offsets[r] + counts[r];
    const long long total = offsets[nrows];
    *total_out = total;
    cudaFree(d_counts);
    if (!d_ids_out) return MPHX_OK;
    int *d_ids = nullptr;
    CK(cudaMalloc(&d_ids, sizeof(int) * (size_t)std::max<long long>(total, 1)));
    CK(cudaMalloc(&d_off, sizeof(long long) * ((size_t)nrows + 1)));
    CK(cudaMemcpyAsync(d_off, offsets.data(), sizeof(long long) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, c->stream));
    if (c->p.dim == 3)
        LAUNCH(c, (k_neighbors_exact<3, 1>), nblk(n), kBlock, n, c->bx, c->by, c->bz, S.type, S.id, S.key, c->cellStart, c->grid,
               cutoff2, structure_only ? 1 : 0, xy_only ? 1 : 0, row_base, (int *)nullptr, d_off, d_ids);
    else
        LAUNCH(c, (k_neighbors_exact<2, 1>), nblk(n), kBlock, n, c->bx, c->by, c->bz, S.type, S.id, S.key, c->cellStart, c->grid,
               cutoff2, structure_only ? 1 : 0, xy_only ? 1 : 0, row_base, (int *)nullptr, d_off, d_ids);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    cudaFree(d_off);
    *d_ids_out = d_ids;
    return MPHX_OK;
}

static int init_solid(Ctx *c)
{
    if (c->ns <= 0) return MPHX_OK;
    const int ns = c->ns;
    int rc;
    // calculateInitialNeighbor (:1497-1644): buckets over InitialPosition, structure particles only.
    // Built on a temporary particle set (the solids in their reference configuration) with the
    // GLOBAL periodic grid, so the lists are complete on every slab of a multi-GPU run.
    Ctx saved = *c; // shallow: pointers / descriptors swapped out below are restored from here
    {
        GridDesc &g = c->grid;
        const mphx_constants &k = c->c;
        g.slab = 0; g.nx = k.cell_count[0]; g.nxg = g.nx; g.xoff = 0; g.mn[0] = c->p.domain_min[0];
        g.ncells = k.cell_counts;
    }
    Particles A{}, B{};
    std::vector<void *> tmp;
    auto talloc = [&](auto **ptr, size_t count) -> int {
        void *q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(**ptr)) != cudaSuccess) return MPHX_ERR_NOMEM;
        tmp.push_back(q);
        *ptr = (std::remove_reference_t<decltype(**ptr)> *)q;
        return 0;
    };
    auto tfree = [&]() { for (void *q : tmp) cudaFree(q); };
    int e = 0;
    for (Particles *p : {&A, &B}) {
        e |= talloc(&p->x, ns); e |= talloc(&p->y, ns); e |= talloc(&p->z, ns);
        e |= talloc(&p->vx, ns); e |= talloc(&p->vy, ns); e |= talloc(&p->vz, ns);
        e |= talloc(&p->type, ns); e |= talloc(&p->id, ns); e |= talloc(&p->key, ns); e |= talloc(&p->pf, (size_t)ns / 2 + 4);
    }
    e |= talloc(&A.ra, (size_t)ns); e |= talloc(&A.rb, (size_t)ns);
    B.ra = A.ra; B.rb = A.rb;
    const size_t ncg = (size_t)c->grid.ncells;
    const int sb_blocks = (int)((ncg + 2 + kScanChunk - 1) / kScanChunk);
    e |= talloc(&c->cellCount, ncg + 2); e |= talloc(&c->cellStart, ncg + 3); e |= talloc(&c->blockSums, (size_t)sb_blocks + 1);
    e |= talloc(&c->slot, ns); e |= talloc(&c->tmpIdx, ns); e |= talloc(&c->where, ns);
    if (e) { tfree(); *c = saved; return MPHX_ERR_NOMEM; }
    c->scan_blocks = sb_blocks;
    c->S = A; c->T = B; c->n = ns; c->slab = false; c->overlap_solid = false; // (restored with the rest of the context)
    LAUNCH(c, k_solid_reference_particles, nblk(ns), kBlock, c->sol, c->S);
    // k_prestep reads a solid's position from the solid arrays: point them at the reference positions
    c->sol.x = saved.sol.x0; c->sol.y = saved.sol.y0; c->sol.z = saved.sol.z0;
    c->sol.slot = nullptr; // (the permute of this temporary set must not overwrite the solids' real slots)
    rc = rebuild_buckets(c, false);
    c->sol = saved.sol;
    if (rc) { tfree(); *c = saved; return rc; }
    std::vector<long long> off;
    int *d_ids = nullptr;
    long long total = 0;
    rc = exact_lists(c, true, c->p.dim == 2, c->sol.sb, ns, off, &d_ids, &total);
    if (rc) { tfree(); *c = saved; return rc; }
    std::vector<int> ids((size_t)std::max<long long>(total, 1));
    {
        cudaError_t ce = cudaMemcpy(ids.data(), d_ids, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost);
        cudaFree(d_ids);
        const long long launches = c->launches;
        tfree();
        *c = saved; // back to the real particle set / grid
        c->launches = launches;
        if (ce != cudaSuccess) { set_last_error("copying the initial lists failed"); return MPHX_ERR_CUDA; }
    }
    if (rc) return rc;
    if (total > 0x7fffffffLL) return MPHX_ERR_UNSUPPORTED;
    std::vector<int> off32((size_t)ns + 1);
    for (int s = 0; s <= ns; ++s) off32[s] = (int)off[s];
    // Store every row in the reference's list order: buckets jCX, jCY, jCZ ascending over the
    // stencil offsets (:1591-1620).  The reference configuration holds at most one solid particle per
    // bucket (lattice at cell spacing); if a bucket ever holds several, ties fall back to id order.
    {
        std::vector<double> hx0(ns), hy0(ns), hz0(ns);
        CK(cudaMemcpy(hx0.data(), c->sol.x0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hy0.data(), c->sol.y0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hz0.data(), c->sol.z0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        GridDesc g = c->grid; // global bucket coordinates (the context's own grid may be one slab)
        g.nx = c->c.cell_count[0]; g.mn[0] = c->p.domain_min[0];
        auto coord = [](double x, double mn, double cw, int n) {
            int v = ((int)std::floor((x - mn) / cw)) % n;
            return (v % n + n) % n;
        };
        std::vector<int> cx(ns), cy(ns), cz(ns);
        for (int s = 0; s < ns; ++s) {
            cx[s] = coord(hx0[s], g.mn[0], g.cellw, g.nx);
            cy[s] = coord(hy0[s], g.mn[1], g.cellw, g.ny);
            cz[s] = g.dim == 3 ? coord(hz0[s], g.mn[2], g.cellw, g.nz) : 0;
        }
        const int R = g.range, span = 2 * R + 1;
        auto off1 = [&](int cj, int ci, int n) {
            int d = cj - ci;
            if (d > R) d -= n;
            if (d < -R) d += n;
            return d + R;
        };
        std::vector<std::pair<long long, int>> row;
        for (int s = 0; s < ns; ++s) {
            row.clear();
            for (int q = off32[s]; q < off32[s + 1]; ++q) {
                const int j = ids[q];
                const long long key = ((long long)off1(cx[j], cx[s], g.nx) * span + off1(cy[j], cy[s], g.ny)) * span +
                                      (g.dim == 3 ? off1(cz[j], cz[s], g.nz) : 0);
                row.emplace_back(key, j);
            }
            std::sort(row.begin(), row.end());
            for (size_t q = 0; q < row.size(); ++q) ids[off32[s] + q] = row[q].second;
        }
    }
    // transpose: roff/rnbr (rows = particles that list s), ascending
    std::vector<int> roff((size_t)ns + 1, 0), rnbr((size_t)std::max<long long>(total, 1));
    for (long long k = 0; k < total; ++k) ++roff[ids[k] + 1];
    for (int s = 0; s < ns; ++s) roff[s + 1] += roff[s];
    {
        std::vector<int> fill(roff.begin(), roff.end() - 1);
        for (int s = 0; s < ns; ++s)
            for (int k = off32[s]; k < off32[s + 1]; ++k) rnbr[fill[ids[k]]++] = s;
    }
    e = 0;
    e |= c->alloc(&c->sol.off, (size_t)ns + 1); e |= c->alloc(&c->sol.nbr, (size_t)total);
    e |= c->alloc(&c->sol.roff, (size_t)ns + 1); e |= c->alloc(&c->sol.rnbr, (size_t)total);
    {
        int maxlen = 0, rmaxlen = 0;
        for (int s = 0; s < ns; ++s) {
            maxlen = std::max(maxlen, off32[s + 1] - off32[s]);
            rmaxlen = std::max(rmaxlen, roff[s + 1] - roff[s]);
        }
        double **pd[] = {&c->sol.d0x, &c->sol.d0y, &c->sol.d0z, &c->sol.w};
        for (double **q : pd) e |= c->alloc(q, (size_t)maxlen * ns);
        double **pr[] = {&c->sol.rd0x, &c->sol.rd0y, &c->sol.rd0z, &c->sol.rw};
        for (double **q : pr) e |= c->alloc(q, (size_t)rmaxlen * ns);
        e |= c->alloc(&c->sol.enbr, (size_t)maxlen * ns); e |= c->alloc(&c->sol.ernbr, (size_t)rmaxlen * ns);
        e |= c->alloc(&c->sol.len, (size_t)ns); e |= c->alloc(&c->sol.rlen, (size_t)ns); e |= c->alloc(&c->sol.rsplit, (size_t)ns);
    }
    if (e) return MPHX_ERR_NOMEM;
    CK(cudaMemcpy(c->sol.off, off32.data(), sizeof(int) * ((size_t)ns + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->sol.nbr, ids.data(), sizeof(int) * (size_t)total, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->sol.roff, roff.data(), sizeof(int) * ((size_t)ns + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->sol.rnbr, rnbr.data(), sizeof(int) * (size_t)total, cudaMemcpyHostToDevice));
    // Lame constants per particle (:2533-2539)
    {
        std::vector<int> ht(ns);
        CK(cudaMemcpy(ht.data(), c->sol.type, sizeof(int) * ns, cudaMemcpyDeviceToHost));
        std::vector<double> lam(ns), mu(ns);
        for (int s = 0; s < ns; ++s) {
            const double E = c->p.young_modulus[ht[s]], v = c->p.poisson_ratio[ht[s]];
            lam[s] = (E * v) / ((1.0 + v) * (1.0 - 2.0 * v));
            mu[s] = E / (2.0 * (1.0 + v));
        }
        CK(cudaMemcpy(c->sol.lam, lam.data(), sizeof(double) * ns, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->sol.mu, mu.data(), sizeof(double) * ns, cudaMemcpyHostToDevice));
    }
    const mphx_constants &k = c->c;
    if (c->p.dim == 3)
        LAUNCH(c, k_solid_pairs<3>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, c->cw_tl);
    else
        LAUNCH(c, k_solid_pairs<2>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, c->cw_tl);
    if (c->p.dim == 3)
        LAUNCH(c, k_solid_normalizer<3>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, c->cw_tl);
    else
        LAUNCH(c, k_solid_normalizer<2>, nblk(ns), kBlock, c->sol, k.domain_width[0], k.domain_width[1], k.domain_width[2], k.radius_p, c->cw_tl);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return MPHX_OK;
}

} // namespace mphx

using namespace mphx;

// =================================================================================================
extern "C" {

int mphx_version(void) { return MPHX_VERSION; }

const char *mphx_strerror(int code)
{
    switch (code) {
    case MPHX_OK: return "ok";
    case MPHX_ERR_INVALID: return "invalid argument or state";
    case MPHX_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case MPHX_ERR_CUDA: return "CUDA error";
    case MPHX_ERR_IO: return "file error";
    case MPHX_ERR_NOMEM: return "out of memory";
    case MPHX_ERR_UNSUPPORTED: return "unsupported configuration";
    case MPHX_ERR_OVERFLOW: return "neighbour capacity exceeded";
    default: return "unknown error";
    }
}

const char *mphx_last_error(void) { return g_last_error.c_str(); }

int mphx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int mphx_abi_sizeof(int which)
{
    switch (which) {
    case 0: return (int)sizeof(mphx_params);
    case 1: return (int)sizeof(mphx_run_control);
    case 2: return (int)sizeof(mphx_constants);
    case 3: return (int)sizeof(mphx_host_views);
    default: return -1;
    }
}

int mphx_create(mphx_ctx **out, const mphx_params *p, int device)
{
    if (!out || !p) return MPHX_ERR_INVALID;
    *out = nullptr;
    if (p->dim != 2 && p->dim != 3) return MPHX_ERR_INVALID;
    if (p->clamp_module < 0 || p->clamp_module > 2) return MPHX_ERR_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_last_error("no CUDA device visible; mphx has no CPU fallback");
        return MPHX_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return MPHX_ERR_INVALID;
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) {
        set_last_error("device is not compute capability 10.x; libmphx.so carries sm_100a code only");
        return MPHX_ERR_NO_DEVICE;
    }
    Ctx *c = new Ctx();
    c->p = *p;
    c->device = device;
    c->time = p->time0;
    for (int t = 0; t < kTypeCount; ++t)
        for (int d = 0; d < 3; ++d) c->wall_center[t][d] = p->wall_center[t][d];
    int rc = setup_constants(c);
    if (rc) { delete c; return rc; }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_last_error("cudaSetDevice/cudaStreamCreate failed");
        delete c;
        return MPHX_ERR_CUDA;
    }
    c->timing = std::getenv("MPHX_TIMING") != nullptr;
    if (const char *e = std::getenv("MPHX_SWEEP_BATCH")) c->sweep_batch = std::max(1, std::atoi(e));
    if (const char *e = std::getenv("MPHX_LIST_CAP")) c->list_cap = std::max(0, std::atoi(e)); // 0: no list, fused sweeps
    if (const char *e = std::getenv("MPHX_FILTER2")) c->filter2 = std::atoi(e) != 0;           // 0: one particle per thread
    for (int k = 0; k < 2; ++k) cudaEventCreate(&c->tev[k]);
    if (const char *e = std::getenv("MPHX_OVERLAP_SOLID")) c->overlap_solid = std::atoi(e) != 0;
    {   // highest priority: the few blocks of a sub-step kernel must get SM slots as pass-2 blocks retire,
        // not after the whole pass-2 grid has been issued
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi) != cudaSuccess) c->side = nullptr;
    }
    cudaEventCreateWithFlags(&c->ev_solid_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_solid_done, cudaEventDisableTiming);
    *out = reinterpret_cast<mphx_ctx *>(c);
    return MPHX_OK;
}

void mphx_destroy(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) if (c->tev[k]) cudaEventDestroy(c->tev[k]);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->ev_solid_ready) cudaEventDestroy(c->ev_solid_ready);
    if (c->ev_solid_done) cudaEventDestroy(c->ev_solid_done);
    for (void *q : c->allocs) cudaFree(q);
    if (c->stream && !c->external_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int mphx_upload(mphx_ctx *ctx, int n, const int *property, const double *position,
                const double *initial_position, const double *velocity)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || n <= 0 || !property || !position || !initial_position || !velocity) return MPHX_ERR_INVALID;
    if (c->uploaded && n != c->n_global) return MPHX_ERR_INVALID; // re-upload must keep the particle count
    if (c->uploaded && c->slab) { set_last_error("re-upload is not supported on a slab context"); return MPHX_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    for (int i = 0; i < n; ++i)
        if (property[i] < 0 || property[i] >= kTypeCount) { set_last_error("particle type outside 0..5"); return MPHX_ERR_INVALID; }
    int r[6];
    mphx_class_ranges(n, property, r);
    // each class must be one contiguous block of the file order, as the reference's range loops
    // (src/main.cpp:909-929, e.g. :2922, :2442) assume
    for (int cls = 0; cls < 3; ++cls)
        for (int i = std::max(r[2 * cls], 0); i < r[2 * cls + 1]; ++i) {
            const int t = property[i];
            const int k = t < 2 ? 0 : t < 4 ? 1 : 2;
            if (k != cls) { set_last_error("particle classes are not contiguous in file order"); return MPHX_ERR_UNSUPPORTED; }
        }
    // slab mode: this context keeps the fluid/wall particles of its own columns and ALL solids
    std::vector<int> ids;
    if (c->slab) {
        const GridDesc &g = c->grid;
        for (int i = 0; i < n; ++i) {
            bool keep = property[i] >= 2 && property[i] < 4;
            if (!keep) {
                int cx = ((int)std::floor((position[3 * (size_t)i] - g.mn0g) / g.cellw)) % g.nxg; // :1671
                cx = (cx % g.nxg + g.nxg) % g.nxg;
                cx -= g.xoff;
                if (cx < 0) cx += g.nxg; else if (cx >= g.nxg) cx -= g.nxg;
                keep = cx >= g.range && cx < g.nx - g.range;
            }
            if (keep) ids.push_back(i);
        }
    }
    const int nloc = c->slab ? (int)ids.size() : n;
    const bool first = !c->uploaded;
    if (first) {
        c->n_global = n;
        c->n = nloc;
        if (!c->slab) c->cap = n;
        if (nloc > c->cap) { set_last_error("slab capacity too small for the initial particle set"); return MPHX_ERR_NOMEM; }
        const size_t cap = (size_t)c->cap;
        std::memcpy(c->ranges, r, sizeof(r));
        c->nf = r[0] >= 0 ? r[1] - r[0] : 0;
        c->ns = r[2] >= 0 ? r[3] - r[2] : 0;
        c->nw = r[4] >= 0 ? r[5] - r[4] : 0;
        int e = 0;
        e |= alloc_particles(c, &c->S, cap);
        e |= alloc_particles(c, &c->T, cap);
        e |= alloc_records(c, &c->S, &c->T, cap);
        {
            int L = c->list_cap;
            if (L < 0) L = c->p.dim == 3 ? 128 : 48;
            c->pl = PairList{};
            if (L > 0) {
                e |= c->alloc(&c->pl.nbr, ((size_t)L + 1) * cap); // + the parking row of overflowed lists
                e |= c->alloc(&c->pl.count, cap);
                e |= c->alloc(&c->pl.flags, 4);
                c->pl.cap = (int)cap; c->pl.L = L;
            }
        }
        e |= c->alloc(&c->cellCount, (size_t)c->grid.ncells + 2);
        e |= c->alloc(&c->cellStart, (size_t)c->grid.ncells + 3);
        c->scan_blocks = (int)(((long long)c->grid.ncells + 2 + kScanChunk - 1) / kScanChunk);
        e |= c->alloc(&c->blockSums, (size_t)c->scan_blocks + 1);
        e |= c->alloc(&c->slot, cap); e |= c->alloc(&c->tmpIdx, cap); e |= c->alloc(&c->where, cap);
        e |= c->alloc(&c->P, cap); e |= c->alloc(&c->volStrain, cap); e |= c->alloc(&c->divP, cap);
        e |= c->alloc(&c->fx, cap); e |= c->alloc(&c->fy, cap); e |= c->alloc(&c->fz, cap);
        e |= c->alloc(&c->ax, cap); e |= c->alloc(&c->ay, cap); e |= c->alloc(&c->az, cap);
        e |= c->alloc(&c->densA, cap); e |= c->alloc(&c->gcx, cap); e |= c->alloc(&c->gcy, cap);
        e |= c->alloc(&c->gcz, cap); e |= c->alloc(&c->PA, cap);
        e |= c->alloc(&c->d_inv_density, kTypeCount);
        e |= c->alloc(&c->d_err, 4);
        if (c->slab) { e |= c->alloc(&c->haloSrc[0], (size_t)c->msg_cap); e |= c->alloc(&c->haloSrc[1], (size_t)c->msg_cap); }
        Solid &so = c->sol;
        so.ns = c->ns; so.sb = c->ns > 0 ? r[2] : 0;
        const size_t ns = (size_t)c->ns;
        double **sv[] = {&so.x, &so.y, &so.z, &so.vx, &so.vy, &so.vz, &so.x0, &so.y0, &so.z0, &so.fx, &so.fy, &so.fz, &so.lam, &so.mu,
                         &so.ux, &so.uy, &so.uz};
        for (double **q : sv) e |= c->alloc(q, ns);
        double **st[] = {&so.Linv, &so.Fm, &so.E, &so.S, &so.Pk};
        for (double **q : st) e |= c->alloc(q, 9 * ns);
        e |= c->alloc(&so.type, ns); e |= c->alloc(&so.slot, ns);
        if (e) return MPHX_ERR_NOMEM;
        CK(cudaMemcpy(c->d_inv_density, c->phys.inv_density, sizeof(double) * kTypeCount, cudaMemcpyHostToDevice));
        CK(cudaMemsetAsync(c->d_err, 0, sizeof(int) * 4, c->stream));
        double *zs[] = {c->P, c->volStrain, c->divP, c->fx, c->fy, c->fz, c->ax, c->ay, c->az, c->densA, c->gcx, c->gcy, c->gcz, c->PA};
        for (double *q : zs) CK(cudaMemsetAsync(q, 0, sizeof(double) * cap, c->stream));
        if (ns > 0)
            for (double **q : st) CK(cudaMemsetAsync(*q, 0, sizeof(double) * 9 * ns, c->stream));
    }
    // stage the host arrays (global, original order), split to SoA on the device
    int *d_t = nullptr, *d_ids = nullptr;
    double *d_x = nullptr, *d_v = nullptr, *d_x0 = nullptr;
    CK(cudaMalloc(&d_t, sizeof(int) * n));
    CK(cudaMalloc(&d_x, sizeof(double) * 3 * (size_t)n));
    CK(cudaMalloc(&d_v, sizeof(double) * 3 * (size_t)n));
    CK(cudaMalloc(&d_x0, sizeof(double) * 3 * (size_t)n));
    CK(cudaMemcpyAsync(d_t, property, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_x, position, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_v, velocity, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_x0, initial_position, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    if (c->slab) {
        CK(cudaMalloc(&d_ids, sizeof(int) * (size_t)std::max(nloc, 1)));
        CK(cudaMemcpyAsync(d_ids, ids.data(), sizeof(int) * (size_t)nloc, cudaMemcpyHostToDevice, c->stream));
    }
    if (nloc > 0) LAUNCH(c, k_upload_split, nblk(nloc), kBlock, nloc, d_ids, d_t, d_x, d_v, c->S, c->sol.slot, c->sol.sb);
    if (c->ns > 0) LAUNCH(c, k_solid_upload, nblk(c->ns), kBlock, c->sol, d_t, d_x, d_x0, d_v);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    cudaFree(d_t); cudaFree(d_x); cudaFree(d_v); cudaFree(d_x0);
    if (d_ids) cudaFree(d_ids);
    c->n = nloc;
    c->uploaded = true;
    c->buckets_valid = false;
    if (!first && c->inited) {
        // state replaced on an initialised context: rebuild the buckets (no wall motion / wrap)
        int rc = rebuild_buckets(c, false);
        if (rc) return rc;
    }
    return MPHX_OK;
}

// replace Position and Velocity of every particle (original order) on an initialised context:
// the per-step host->device path of a caller that owns the state on the host
int mphx_upload_state(mphx_ctx *ctx, const double *position, const double *velocity)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !position || !velocity) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    const size_t N = (size_t)c->n_global;
    if (!c->stage3a && c->alloc(&c->stage3a, 3 * N)) return MPHX_ERR_NOMEM;
    if (!c->stage3b && c->alloc(&c->stage3b, 3 * N)) return MPHX_ERR_NOMEM;
    CK(cudaMemcpyAsync(c->stage3a, position, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->stage3b, velocity, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_upload_state, nblk(c->n), kBlock, c->n, c->S, c->sol, c->stage3a, c->stage3b);
    c->buckets_valid = false;
    CK(cudaGetLastError());
    return MPHX_OK;
}

int mphx_init(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->uploaded) return MPHX_ERR_INVALID;
    if (c->inited) return MPHX_OK;
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = init_solid(c))) return rc; // calculateInitialNeighbor, Lame, Normalizer
    // first calculateNeighbor + density sums on the initial positions (:565-568): gives
    // NeighborCount / PressureP for the `output.vtk` written before the loop (:572)
    if (!c->slab) { // (a slab needs its neighbours' halos first: done by the first mphx_slab_* step)
        if ((rc = rebuild_buckets(c, false))) return rc;
        if ((rc = run_pass1(c))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    c->inited = true;
    return MPHX_OK;
}

int mphx_get_constants(const mphx_ctx *ctx, mphx_constants *k)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !k) return MPHX_ERR_INVALID;
    *k = c->c;
    return MPHX_OK;
}

int mphx_step(mphx_ctx *ctx, int nsteps)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || nsteps < 0) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    for (int s = 0; s < nsteps; ++s) {
        int rc = one_step(c, false);
        if (rc) return rc;
    }
    return MPHX_OK;
}

int mphx_timed_steps(mphx_ctx *ctx, int nsteps, double *elapsed_ms)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || nsteps < 0 || !elapsed_ms) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->tev[0], c->stream));
    for (int s = 0; s < nsteps; ++s) {
        int rc = one_step(c, false);
        if (rc) return rc;
    }
    { int jrc = join_solids(c); if (jrc) return jrc; } // the timed region ends when the last sub-steps have finished
    CK(cudaEventRecord(c->tev[1], c->stream));
    CK(cudaEventSynchronize(c->tev[1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->tev[0], c->tev[1]));
    *elapsed_ms = ms;
    CK(cudaGetLastError());
    return MPHX_OK;
}

int mphx_set_timing(mphx_ctx *ctx, int on)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return MPHX_ERR_INVALID;
    cudaSetDevice(c->device);
    timer_resolve(c);
    c->timing = on != 0;
    if (on) for (int k = 0; k < 5; ++k) c->ms[k] = 0.0;
    return MPHX_OK;
}

int mphx_step_fluid_only(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    return one_step(c, true);
}

int mphx_sync(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return MPHX_OK;
}

double mphx_time(const mphx_ctx *ctx) { return ctx ? reinterpret_cast<const Ctx *>(ctx)->time : 0.0; }

int mphx_set_time(mphx_ctx *ctx, double t)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return MPHX_ERR_INVALID;
    c->time = t;
    return MPHX_OK;
}

int mphx_download(mphx_ctx *ctx, const mphx_host_views *v)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !v || !c->uploaded) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    // Arrays are in ORIGINAL particle order and sized for the whole case.  A slab context fills the
    // entries of the particles it owns (solids: slab 0 only) and zeros elsewhere, so the caller can
    // combine the slabs with a plain sum.
    const int n = c->n, ns = c->ns;
    const size_t N = (size_t)c->n_global;
    if (!c->stage3a && c->alloc(&c->stage3a, 3 * N)) return MPHX_ERR_NOMEM;
    if (!c->stage1 && c->alloc(&c->stage1, N)) return MPHX_ERR_NOMEM;
    if (!c->stagei && c->alloc(&c->stagei, N)) return MPHX_ERR_NOMEM;
    if (!c->mask && c->alloc(&c->mask, (size_t)c->cap)) return MPHX_ERR_NOMEM;
    if (!c->mask2 && c->alloc(&c->mask2, (size_t)c->cap)) return MPHX_ERR_NOMEM;
    if (!c->tmpi && c->alloc(&c->tmpi, (size_t)c->cap)) return MPHX_ERR_NOMEM;
    double *d3 = c->stage3a, *d1 = c->stage1, *d9 = c->stage9;
    int *di = c->stagei;
    const Particles &S = c->S;
    const Solid &so = c->sol;
    const bool report_solids = !c->slab || c->rank == 0;
    // mask: per-particle state (solids: one reporting slab, values come from the replicated solid arrays);
    // mask2: fields evaluated per bucket sweep (PressureP, VolStrainP, CellIndex, ...), reported for a solid
    // by the slab that owns its current column
    if (n > 0) LAUNCH(c, k_owned_mask, nblk(n), kBlock, n, S, c->grid, report_solids ? 1 : 0, c->mask);
    if (n > 0) LAUNCH(c, k_owned_mask, nblk(n), kBlock, n, S, c->grid, 2, c->mask2);
    int rc = MPHX_OK;
    auto vec3 = [&](double *host, const double *a, const double *b, const double *cc, const double *sa, const double *sb_, const double *sc) -> int {
        if (!host) return MPHX_OK;
        if (c->slab) CK(cudaMemsetAsync(d3, 0, sizeof(double) * 3 * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_vec3, nblk(n), kBlock, n, S.id, c->mask, a, b, cc, d3);
        if (ns > 0 && sa && report_solids) LAUNCH(c, k_solid_vec3_to_orig, nblk(ns), kBlock, so, sa, sb_, sc, d3);
        CK(cudaMemcpyAsync(host, d3, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MPHX_OK;
    };
    auto scal = [&](double *host, const double *a) -> int {
        if (!host) return MPHX_OK;
        if (c->slab) CK(cudaMemsetAsync(d1, 0, sizeof(double) * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_scalar, nblk(n), kBlock, n, S.id, c->mask2, a, d1);
        CK(cudaMemcpyAsync(host, d1, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MPHX_OK;
    };
    auto ints = [&](int *host, const int *a, const int *msk) -> int {
        if (!host) return MPHX_OK;
        if (c->slab) CK(cudaMemsetAsync(di, 0, sizeof(int) * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_int, nblk(n), kBlock, n, S.id, msk, a, di);
        CK(cudaMemcpyAsync(host, di, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MPHX_OK;
    };
    auto tens = [&](double *host, const double *M) -> int {
        if (!host) return MPHX_OK;
        if (!d9) {
            if (c->alloc(&c->stage9, 9 * N)) return MPHX_ERR_NOMEM;
            d9 = c->stage9;
        }
        CK(cudaMemsetAsync(d9, 0, sizeof(double) * 9 * N, c->stream));
        if (ns > 0 && report_solids) LAUNCH(c, k_solid_tensor_to_orig, nblk(ns), kBlock, so, M, d9);
        CK(cudaMemcpyAsync(host, d9, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MPHX_OK;
    };
    auto solid_scal = [&](double *host, const double *a) -> int {
        if (!host) return MPHX_OK;
        CK(cudaMemsetAsync(d1, 0, sizeof(double) * N, c->stream));
        if (ns > 0 && report_solids) LAUNCH(c, k_solid_scalar_to_orig, nblk(ns), kBlock, so, a, d1);
        CK(cudaMemcpyAsync(host, d1, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MPHX_OK;
    };
    do {
        if (v->property) {
            if (n > 0) LAUNCH(c, k_real_types, nblk(n), kBlock, n, S, c->tmpi);
            if ((rc = ints(v->property, c->tmpi, c->mask))) break;
        }
        if ((rc = vec3(v->position, S.x, S.y, S.z, so.x, so.y, so.z))) break;
        if ((rc = vec3(v->velocity, S.vx, S.vy, S.vz, so.vx, so.vy, so.vz))) break;
        if ((rc = vec3(v->force, c->fx, c->fy, c->fz, so.fx, so.fy, so.fz))) break;
        if ((rc = vec3(v->acceleration, c->ax, c->ay, c->az, nullptr, nullptr, nullptr))) break;
        if ((rc = scal(v->pressure_p, c->P))) break;
        if ((rc = scal(v->vol_strain_p, c->volStrain))) break;
        if ((rc = scal(v->divergence_p, c->divP))) break;
        if ((rc = scal(v->density_a, c->densA))) break;
        if ((rc = vec3(v->gravity_center, c->gcx, c->gcy, c->gcz, nullptr, nullptr, nullptr))) break;
        if ((rc = scal(v->pressure_a, c->PA))) break;
        if (v->cell_index) {
            if (n > 0) LAUNCH(c, k_global_keys, nblk(n), kBlock, n, S, c->grid, c->tmpi);
            if ((rc = ints(v->cell_index, c->tmpi, c->mask2))) break;
        }
        if (v->neighbor_count) {
            if (!c->inited) { rc = MPHX_ERR_INVALID; break; }
            if (!c->buckets_valid) {
                if (c->slab) { rc = MPHX_ERR_INVALID; break; }
                if ((rc = rebuild_buckets(c, false))) break;
            }
            std::vector<long long> off;
            long long total = 0;
            if ((rc = exact_lists(c, false, false, 0, (int)N, off, nullptr, &total))) break;
            for (size_t i = 0; i < N; ++i) v->neighbor_count[i] = (int)(off[i + 1] - off[i]);
        }
        if (v->initial_structure_neighbor_count) {
            CK(cudaMemsetAsync(di, 0, sizeof(int) * N, c->stream));
            if (ns > 0 && so.off && report_solids) LAUNCH(c, k_solid_rowlen_to_orig, nblk(ns), kBlock, so, di);
            CK(cudaMemcpyAsync(v->initial_structure_neighbor_count, di, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        if ((rc = tens(v->normalizer, so.Linv))) break;
        if ((rc = tens(v->deform_gradient, so.Fm))) break;
        if ((rc = tens(v->strain, so.E))) break;
        if ((rc = tens(v->stress, so.S))) break;
        if ((rc = solid_scal(v->lambda_lames, so.lam))) break;
        if ((rc = solid_scal(v->mu_lames, so.mu))) break;
    } while (0);
    if (rc == MPHX_OK) { CK(cudaGetLastError()); }
    return rc;
}

// ---- compact I/O of the particles a context owns ------------------------------------------------------
// (single context: every particle; slab context: the fluid/wall particles of its columns plus ALL of the
// replicated solids).  Rows come in the context's current slot order; mphx_upload_owned takes the rows of
// the last mphx_download_owned back (same order, possibly modified values) -- the per-step host<->device
// path of a distributed caller, moving 1/nranks of the state per rank.
static int owned_prepare(Ctx *c)
{
    const size_t cap = (size_t)c->cap;
    if (!c->mask && c->alloc(&c->mask, cap)) return MPHX_ERR_NOMEM;
    if (!c->own_scan) {
        c->own_blocks = (int)((cap + 1 + kScanChunk - 1) / kScanChunk);
        int e = 0;
        e |= c->alloc(&c->own_scan, cap + 2); e |= c->alloc(&c->own_sums, (size_t)c->own_blocks + 1);
        e |= c->alloc(&c->own_slot, cap); e |= c->alloc(&c->own_ids, cap);
        e |= c->alloc(&c->own_x, 3 * cap); e |= c->alloc(&c->own_v, 3 * cap);
        if (e) return MPHX_ERR_NOMEM;
    }
    return MPHX_OK;
}

int mphx_download_owned(mphx_ctx *ctx, int capacity, int *ids, double *position, double *velocity, int *count)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !ids || !position || !velocity || !count) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    int rc = owned_prepare(c);
    if (rc) return rc;
    const int n = c->n;
    LAUNCH(c, k_owned_mask, nblk(n), kBlock, n, c->S, c->grid, 1, c->mask);
    const int nb = (int)(((long long)n + kScanChunk - 1) / kScanChunk);
    LAUNCH(c, k_scan_reduce, nb, kScanThreads, c->mask, n, c->own_sums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, c->own_sums, nb);
    LAUNCH(c, k_scan_apply, nb, kScanThreads, c->mask, n, c->own_sums, c->own_scan);
    int total = 0;
    CK(cudaMemcpyAsync(&total, c->own_scan + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LAUNCH(c, k_compact_owned, nblk(n), kBlock, n, c->S, c->sol, c->mask, c->own_scan, c->own_slot, c->own_ids, c->own_x, c->own_v);
    CK(cudaStreamSynchronize(c->stream));
    *count = total;
    if (total > capacity) { set_last_error("mphx_download_owned: capacity too small"); return MPHX_ERR_OVERFLOW; }
    CK(cudaMemcpyAsync(ids, c->own_ids, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(position, c->own_x, sizeof(double) * 3 * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(velocity, c->own_v, sizeof(double) * 3 * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->own_count = total;
    c->own_epoch = c->steps_done;
    CK(cudaGetLastError());
    return MPHX_OK;
}

int mphx_upload_owned(mphx_ctx *ctx, int count, const int *ids, const double *position, const double *velocity)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !ids || !position || !velocity) return MPHX_ERR_INVALID;
    if (c->own_epoch != c->steps_done || count != c->own_count) {
        set_last_error("mphx_upload_owned: rows must be those of the last mphx_download_owned (no step in between)");
        return MPHX_ERR_INVALID;
    }
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    CK(cudaMemcpyAsync(c->own_ids, ids, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->own_x, position, sizeof(double) * 3 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->own_v, velocity, sizeof(double) * 3 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    LAUNCH(c, k_scatter_owned, nblk(count), kBlock, count, c->S, c->sol, c->own_slot, c->own_ids, c->own_x, c->own_v, c->d_err);
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (err) { set_last_error("mphx_upload_owned: ids do not match the rows of the last download"); return MPHX_ERR_INVALID; }
    c->buckets_valid = false;
    CK(cudaGetLastError());
    return MPHX_OK;
}

int mphx_debug_neighbors(mphx_ctx *ctx, long long *offsets, int *ids, long long cap)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !offsets) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    std::vector<long long> off;
    long long total = 0;
    int *d_ids = nullptr;
    if (!c->buckets_valid) return MPHX_ERR_INVALID;
    int rc = exact_lists(c, false, false, 0, c->n_global, off, ids ? &d_ids : nullptr, &total);
    if (rc) return rc;
    std::memcpy(offsets, off.data(), sizeof(long long) * ((size_t)c->n_global + 1));
    if (ids) {
        if (cap < total) { cudaFree(d_ids); return MPHX_ERR_OVERFLOW; }
        CK(cudaMemcpy(ids, d_ids, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost));
        cudaFree(d_ids);
    }
    return MPHX_OK;
}

int mphx_debug_initial_structure_neighbors(mphx_ctx *ctx, long long *offsets, int *ids, long long cap)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !offsets) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    const int n = c->n_global, ns = c->ns, sb = c->sol.sb;
    std::vector<int> off32((size_t)ns + 1, 0);
    if (ns > 0) CK(cudaMemcpy(off32.data(), c->sol.off, sizeof(int) * ((size_t)ns + 1), cudaMemcpyDeviceToHost));
    for (int i = 0; i <= n; ++i) {
        const int s = i - sb;
        offsets[i] = (ns > 0 && s >= 0) ? off32[std::min(s, ns)] : 0;
    }
    const long long total = ns > 0 ? off32[ns] : 0;
    if (ids) {
        if (cap < total) return MPHX_ERR_OVERFLOW;
        if (total > 0) {
            CK(cudaMemcpy(ids, c->sol.nbr, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost));
            for (long long k = 0; k < total; ++k) ids[k] += sb;
            for (int q = 0; q < ns; ++q) std::sort(ids + off32[q], ids + off32[q + 1]);
        }
    }
    return MPHX_OK;
}

int mphx_get_timers(mphx_ctx *ctx, double ms[4])
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !ms) return MPHX_ERR_INVALID;
    cudaSetDevice(c->device);
    timer_resolve(c);
    ms[0] = c->ms[0]; ms[1] = c->ms[1] + c->ms[2]; ms[2] = c->ms[3]; ms[3] = c->ms[4];
    return MPHX_OK;
}

int mphx_get_kernel_timers(mphx_ctx *ctx, double ms[5])
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !ms) return MPHX_ERR_INVALID;
    cudaSetDevice(c->device);
    timer_resolve(c);
    for (int i = 0; i < 5; ++i) ms[i] = c->ms[i];
    return MPHX_OK;
}

int mphx_set_overlap(mphx_ctx *ctx, int on)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    c->overlap_solid = on != 0;
    return MPHX_OK;
}

int mphx_join(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    return join_solids(c);
}

long long mphx_launch_count(const mphx_ctx *ctx) { return ctx ? reinterpret_cast<const Ctx *>(ctx)->launches : 0; }

double mphx_algorithmic_bytes_per_step(const mphx_ctx *ctx)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c) return 0.0;
    const int nsub = (int)(c->p.dt / c->p.elastic_dt + 0.5);
    return 368.0 * c->nf + 260.0 * c->nw + (344.0 + 384.0 * nsub) * c->ns;
}

#include "slab.inc"

} // extern "C"
