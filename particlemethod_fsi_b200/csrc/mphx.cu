// mphx.cu -- context, step orchestration and the extern-"C" layer (include/mphx.h).
//
// One context drives one B200 (or one x-slab of a multi-GPU run).  All state lives on the device as
// cell-sorted SoA; the host only keeps scalars (Time, wall centres) and enqueues kernels on the context's
// stream.  Everything a step decides -- rebuild or reuse of the candidate list, slot counts, exchange counts
// -- is decided ON THE DEVICE (struct Ctl), so a step is a fixed sequence of launches without a single
// device->host read-back; slabs exchange particles through peer-mapped mailboxes (see kernels.cuh).
// There is no CPU fallback: every compute entry point fails with MPHX_ERR_NO_DEVICE / MPHX_ERR_CUDA
// when no sm_100 device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "sweep.cuh"
#include "sweep_pair.cuh"
#include "virial.cuh"
#include "brick.cuh"
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
    // Particle slots: the exact number held lives on the device (ctl->n: owned + ghosts + all solids in slab
    // mode, changing with every rebuild); the host only knows a launch bound.
    int nmax = 0;     // launch bound on the slots held (single context: n; slab: capacity)
    int n_global = 0; // particles of the whole case
    int cap = 0;      // allocated slots
    int nf = 0, ns = 0, nw = 0; // class counts of the whole case
    int ranges[6];
    Ctl *ctl = nullptr;
    // slab mode (multi-GPU): see slab.inc
    bool slab = false, connected = false;
    int rank = 0, nranks = 1, col_lo = 0, col_hi = 0, msg_cap = 0;
    int pending_lo = -1, pending_hi = -1; // re-cut requested (mphx_slab_recut): takes effect at the start of the next step
    size_t cells_cap = 0;                 // bucket arrays hold this many buckets (+ service buckets)
    void *mailbox = nullptr;
    size_t mailbox_bytes = 0;
    Mailbox mine{};
    Peers peers{};
    // split solid sub-steps: this rank advances the solids [s_lo, s_hi) and stores P / u / the final state into the peers' arrays
    // sub-step kernels with a team of lanes per solid (kernels.cuh); 0: one thread per solid
    bool solid_team = true, team_ok = false, solid_team_always = false;
    int sol_maxlen = 0, sol_rmaxlen = 0;
    bool split_substeps = true, ring_inkernel_wait = false;
    int s_lo = 0, s_hi = 0;
    SolidRing ring{};
    void *peer_base[kMaxRanks] = {};
    bool peer_ipc[kMaxRanks] = {};
    double *stage_mig[2] = {nullptr, nullptr}, *stage_halo[2] = {nullptr, nullptr};
    int *haloSrc[2] = {nullptr, nullptr}, *haloSlot[2] = {nullptr, nullptr}, *ghostSlot[2] = {nullptr, nullptr}, *own_sol = nullptr;
    int *where = nullptr;
    bool external_stream = false;
    bool uploaded = false, inited = false, surface_tension = false;
    int solid_tuples = 0; // entries of the solids' pair-data dictionary (0: raw arrays in use)
    bool solid_multi_occupancy = false; // several solids share a bucket of the reference configuration (see init_solid)
    unsigned long long epoch = 0; // exchange epoch: one per bucket stage, identical on every rank (flags carry it)
    double time = 0.0;
    double wall_center[kTypeCount][3];
    long long launches = 0;
    long long steps_done = 0;

    GridDesc grid;
    Phys phys;
    Particles S{}, T{}; // S: current cell-sorted state, T: scratch (permute target / pre-step positions)
    double *bx = nullptr, *by = nullptr, *bz = nullptr; // positions of the last pass 1 (the reference's NeighborCount refers to them)
    double *ancx = nullptr, *ancy = nullptr, *ancz = nullptr; // positions the current candidate list was built on
    int *cellCount = nullptr, *cellStart = nullptr, *slot = nullptr, *tmpIdx = nullptr, *blockSums = nullptr;
    int scan_blocks = 0;
    double *P = nullptr, *volStrain = nullptr, *divP = nullptr;
    double *densA = nullptr, *gcx = nullptr, *gcy = nullptr, *gcz = nullptr, *PA = nullptr;
    double *fx = nullptr, *fy = nullptr, *fz = nullptr, *ax = nullptr, *ay = nullptr, *az = nullptr;
    double *virial = nullptr; // [10][cap]: VirialStressAtParticle (9 planes) + VirialPressureAtParticle, lazy
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
    int sweep_batch = 12;  // stencil columns per filter/drain batch (3D)
    PairList pl{};         // candidate list (nbr == nullptr: disabled, both passes walk the buckets)
    int list_cap = -1;     // list slots per particle (-1: default by dimension, 0: no list)
    bool filter2 = true;   // build the lists two particles per thread (k_filter2, sweep_pair.cuh)
    // pass 1 with the neighbourhood staged in shared memory, one block per brick of buckets (brick.cuh; MPHX_BRICK=1)
    bool brick = false;
    unsigned short *lnbr = nullptr;      // the candidate list re-indexed to brick-local positions, in brick order
    int *brick_nown = nullptr, *brick_base = nullptr, *brick_sums = nullptr;
    int brick_scan_blocks = 0;
    unsigned char *brick_ok = nullptr, *in_brick = nullptr;
    int nbricks = 0;
    // candidate-list reuse (internal Verlet skin): the list is built with radius + skin and serves until a
    // particle has moved skin/2 from its build position (decided on the device, k_decide)
    bool list_reuse = true;
    double skin = 0.25;    // in particle spacings
    // the solid sub-steps only need the solids' share of pass 2: they run on a second stream while the
    // fluid's share (FP64 / issue bound; the sub-steps are HBM bound) is still being computed
    bool overlap_solid = true;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_solid_ready = nullptr, ev_solid_done = nullptr, ev_u_ready = nullptr;
    bool early_pass1 = true, early_done = false; // pass 1 of the first sub-step right after the solids' pre-step (second stream)
    bool solids_pending = false;  // sub-steps enqueued on `side`, not yet joined by the main stream
    std::vector<void *> allocs;

    bool tracing = false; // device-side timeline (mphx_trace_enable)
    unsigned long long *trace_buf = nullptr;
    int trace_cap_alloc = 0;
    // phase timers (src/main.cpp:695-700 split)
    bool timing = false;
    std::vector<cudaEvent_t> ev, ev_pool;
    double ms[5] = {0, 0, 0, 0, 0}; // rebuild, filter, pass 1, pass 2, solid sub-steps
    cudaEvent_t tev[2] = {nullptr, nullptr};
    cudaEvent_t side_ev[2] = {nullptr, nullptr}; // sub-steps on the side stream (timing)
    std::vector<cudaEvent_t> side_marks;
    double side_ms = 0.0;
    double virial_ms = 0.0; // device time spent in calculateVirialStressAtParticle (the reference's "virial calculation" timer)

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
#define LAUNCH(ctx, kernel, grid, block, ...) LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, __VA_ARGS__)

// temporary device allocations of one call: freed on every exit path
struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() { for (void *q : ptrs) cudaFree(q); }
    template <class Tp> int get(Tp **ptr, size_t count)
    {
        void *q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(Tp)) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("cudaMalloc (scratch) failed");
            return MPHX_ERR_NOMEM;
        }
        ptrs.push_back(q);
        *ptr = (Tp *)q;
        return MPHX_OK;
    }
};

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

// largest kernel radius either pass uses: they share one candidate list
static double sweep_radius(const Ctx *c)
{
    const mphx_constants &k = c->c;
    double rmax = std::max(k.radius_p, k.radius_v);
    if (c->surface_tension) rmax = std::max(rmax, k.radius_a);
    return rmax;
}
// the Verlet skin actually used (metres): never wider than what the bucket stencil (range * CellWidth) covers
static double skin_length(const Ctx *c)
{
    const double room = (double)c->grid.range * c->grid.cellw - sweep_radius(c);
    return std::max(0.0, std::min(c->skin * c->p.particle_spacing, 0.9 * room));
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
    // the stencil must reach the reference's list cut-off MaxRadius+MARGIN (:116,:1765) and the candidate
    // list's own radius (largest kernel radius + skin)
    const double cutoff = std::max(k.max_radius + 0.1 * p.particle_spacing, sweep_radius(c) + skin_length(c));
    rc = build_stencil(g, cutoff);
    if (rc) { set_last_error("stencil too large (radius ratio too big)"); return rc; }
    return MPHX_OK;
}

// The solid sub-steps run on a second stream and are joined as late as possible: they overlap the fluid's
// share of pass 2 and the fluid/wall share of the next pre-step.
static bool defer_solids(const Ctx *c) { return c->overlap_solid && c->ns > 0 && c->side != nullptr && c->pl.nbr != nullptr; }
static int join_solids(Ctx *c)
{
    if (c->solids_pending) {
        CK(cudaStreamWaitEvent(c->stream, c->ev_solid_done, 0));
        c->solids_pending = false;
    }
    return MPHX_OK;
}

// the exact slot count (device -> host; synchronises the context's stream)
static int read_n(Ctx *c, int *n)
{
    CK(cudaMemcpyAsync(n, &c->ctl->n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MPHX_OK;
}
static int request_rebuild(Ctx *c)
{
    const int one = 1;
    CK(cudaMemcpyAsync(&c->ctl->force, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    return MPHX_OK;
}

// squared cut-off (bucket units) of the fp32 candidate filter: the exact cut-off plus a margin that
// covers the rounding of the fp32 bucket coordinates (<= 2 ulp at the largest coordinate, per axis)
// and of the fp32 distance arithmetic, so the filter is a superset of the fp64 predicate.
static float filter_radius2(const Ctx *c, double rmax)
{
    const GridDesc &g = c->grid;
    const double R = rmax / g.cellw;
    const double big = (double)std::max(std::max(std::max(g.nx, g.nxg), g.ny), std::max(g.nz, 4)) + 2.0 * g.range + 2.0;
    const double delta = std::ldexp(big, -22);           // 2 ulp_f32(big) per coordinate, both particles
    const double margin = 2.0 * std::sqrt(3.0) * (R + 1.0) * (2.0 * delta) + 3.0 * (2.0 * delta) * (2.0 * delta) +
                          8.0 * std::ldexp((R + 1.0) * (R + 1.0), -23) + 1e-6;
    return (float)(R * R + margin) * (1.0f + 1e-6f);
}

static bool walls_move(const Ctx *c)
{
    if (c->nw > 0 && c->p.wall_module == MPHX_WALL_ROLLING) return true;
    if (c->nw <= 0 || !(c->time < 0.2)) return false; // Q7 (:3037)
    for (int t = 4; t < kTypeCount; ++t)
        for (int d = 0; d < 3; ++d)
            if (c->p.wall_velocity[t][d] != 0.0 || c->p.wall_omega[t][d] != 0.0) return true;
    return false;
}

static DecideArgs decide_args(const Ctx *c)
{
    DecideArgs a{};
    const double skin = skin_length(c);
    // moving walls travel |V| dt + |omega| r dt per step without entering the fluid's displacement bound: rebuild
    // every step while they move (t < 0.2, Q7)
    a.reuse_enabled = (c->list_reuse && c->pl.nbr != nullptr && skin > 0.0 && !walls_move(c)) ? 1 : 0;
    a.half_skin2 = (float)(0.25 * skin * skin * (1.0 - 1e-6));
    a.filt2_plain = filter_radius2(c, sweep_radius(c));
    a.filt2_skin = filter_radius2(c, sweep_radius(c) + skin);
    return a;
}

static void timer_mark(Ctx *c);

// CUDA loads kernels lazily, at their first launch, and that load may wait for every kernel running on the device.
// A slab step parks one-warp kernels that spin on a peer's flag; a first-time load issued behind such a kernel would
// stall until the wait times out.  So every kernel of the stepping path is loaded when a context is created.
template <class K> static void preload(K kernel)
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, kernel);
}
static void preload_kernels(int dim)
{
    preload(k_mark); preload(k_prestep); preload(k_need); preload(k_decide); preload(k_vote); preload(k_wait<0>); preload(k_wait<1>);
    preload(k_count); preload(k_wait_decide); preload(k_halo_resend); preload(k_push); preload(k_push_scalar); preload(k_unpack_particles);
    preload(k_advance_n); preload(k_halo_pack); preload(k_unpack_refresh); preload(k_slab_slots);
    preload(k_unpack_scalar); preload(k_solid_owned_list); preload(k_solid_publish_P); preload(k_solid_spread_P);
    preload(k_solid_publish_V); preload(k_solid_apply_update); preload(k_scan_reduce); preload(k_scan_top); preload(k_scan_apply);
    preload(k_scatter_index); preload(k_permute); preload(k_set_n); preload(k_brick_localize); preload(k_brick_count); preload(k_brick_pass1<3>);
#define PRELOAD_DIM(D)                                                                                                   \
    preload(k_filter<D>); preload(k_filter2<D, false>); preload(k_filter2<D, true>);                                     \
    preload(k_pass1_v3<D, false, false>); preload(k_pass1_v3<D, false, true>); preload(k_pass1_v3<D, true, false>);      \
    preload(k_pass1_v3<D, true, true>); preload(k_pass2_v3<D, false, false>); preload(k_pass2_v3<D, false, true>);       \
    preload(k_pass2_v3<D, true, false>); preload(k_pass2_v3<D, true, true>);                                             \
    preload(k_solid_pass1<D, false, false, false>); preload(k_solid_pass2<D, false, false, false>);                      \
    preload(k_solid_pass1<D, true, false, false>); preload(k_solid_pass2<D, true, false, false>);                        \
    preload(k_solid_pass1<D, false, true, true>); preload(k_solid_pass2<D, false, true, true>);                          \
    preload(k_solid_pass1<D, true, true, true>); preload(k_solid_pass2<D, true, true, true>);                            \
    preload(k_solid_pass1<D, true, false, true>); preload(k_solid_pass2<D, true, false, true>);                          \
    preload(k_solid_pass1_team<D, false>); preload(k_solid_pass1_team<D, true>); preload(k_solid_pass2_team<D, false>);  \
    preload(k_solid_pass2_team<D, true>)
    if (dim == 3) { PRELOAD_DIM(3); } else { PRELOAD_DIM(2); }
#undef PRELOAD_DIM
    cudaGetLastError();
}

// ---- exchange between slabs: all device-side (see kernels.cuh "Exchange between slabs") --------------------
constexpr int kPushBlocks = 64, kPushThreads = 256;
enum { kPushMigL = 0, kPushMigR, kPushHaloL, kPushHaloR, kPushPL, kPushPR, kPushSolP, kPushSolV };

// grids of the kernels that only work on rebuild steps (or on a message): grid-stride, so that a step that reuses its
// list does not pay for tens of thousands of blocks that return at once
static int small_grid(long long n, int threads = kBlock) { return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, 148 * 16)); }

static PushPair push_pair(Ctx *c, bool halo)
{
    // what I send "to the left" arrives in the left neighbour's mailbox as "from the right" (side 1), and vice versa
    Ctl *ctl = c->ctl;
    Mailbox &L = c->peers.left, &R = c->peers.right;
    PushPair pp{};
    double **stage = halo ? c->stage_halo : c->stage_mig;
    int *cnt = halo ? ctl->halo_cnt : ctl->mig_cnt;
    pp.src[0] = stage[0]; pp.src[1] = stage[1];
    pp.cnt[0] = cnt + 0; pp.cnt[1] = cnt + 1;
    pp.dst[0] = halo ? L.halo[1] : L.mig[1]; pp.dst[1] = halo ? R.halo[0] : R.mig[0];
    pp.dst_cnt[0] = (halo ? L.cnt_halo : L.cnt_mig) + 1; pp.dst_cnt[1] = (halo ? R.cnt_halo : R.cnt_mig) + 0;
    pp.dst_flag[0] = (halo ? L.fhalo : L.fmig) + 1; pp.dst_flag[1] = (halo ? R.fhalo : R.fmig) + 0;
    return pp;
}
// migration: rebuild steps only (nothing is sent, nothing is waited for on a step that reuses its list);
// halo: rebuild steps push the freshly packed staging buffers, reuse steps pack the same particles straight into the
// neighbours' mailboxes (k_halo_resend); one wait for both neighbours
static int exchange_particles(Ctx *c, bool halo)
{
    Ctl *ctl = c->ctl;
    const PushPair pp = push_pair(c, halo);
    const dim3 grid(kPushBlocks, 2);
    k_push<<<grid, kPushThreads, 0, c->stream>>>(ctl, c->epoch, halo ? kPushHaloL : kPushMigL, pp, kMsgDoubles, c->msg_cap, 1);
    ++c->launches;
    LAUNCH(c, k_wait<0>, 1, 32, ctl, c->epoch, halo ? c->mine.fhalo : c->mine.fmig, 2, halo ? kWaitHalo : kWaitMig, halo ? 0 : 1);
    CK(cudaGetLastError());
    return MPHX_OK;
}

// ---- bucket rebuild / refresh: K0 pre-step, decision, K1 count (+ slab exchange), K2 scan, K3 scatter, K4 permute ---
// motion: the step's calculateWall / calculatePeriodicBoundary (false: buckets over the positions as they are)
static int run_solid_substeps(Ctx *c, cudaStream_t strm, int part);
static bool solid_split(const Ctx *c);
// early_pass1: the step's sub-steps follow (one_step): their first pass 1 starts on the second stream once the solids' pre-step is done
static int stage_build(Ctx *c, bool motion, bool early_pass1 = false)
{
    Ctl *ctl = c->ctl;
    const int nb = nblk(c->nmax);
    ++c->epoch;
    WallMotion wm;
    std::memset(&wm, 0, sizeof(wm));
    wm.dt = c->p.dt;
    const bool rolling = c->p.wall_module == MPHX_WALL_ROLLING;
    wm.active = (motion && c->nw > 0 && (rolling || c->time < 0.2)) ? 1 : 0; // Q7 (:3037); the Rolling variant has no time gate
    if (wm.active && rolling) {
        // `#define Rolling` (:2958-3031): the walls turn about z by theta(t) - theta(t - Dt), theta = 2 deg sin(2 pi t / 1.646 s);
        // v = (0, 0, dtheta/dt) x r_rot, x = r_rot + centre (no translation term) -- the generic wall formula of k_prestep
        // with R = Rz(delta theta), omega = (0, 0, dtheta/dt), V = 0 gives the same bits (the extra terms are exact zeros)
        const double max_angle = 2.0 * M_PI / 180.0, period = 1.646;
        const double omega_t = 2.0 * M_PI / period;
        const double theta = max_angle * std::sin(omega_t * c->time);
        const double dtheta_dt = max_angle * omega_t * std::cos(omega_t * c->time);
        const double theta_prev = max_angle * std::sin(omega_t * (c->time - c->p.dt));
        const double delta_theta = theta - theta_prev;
        const double cosD = std::cos(delta_theta), sinD = std::sin(delta_theta);
        for (int t = 0; t < kTypeCount; ++t) {
            for (int d = 0; d < 3; ++d) { wm.center[t][d] = c->wall_center[t][d]; wm.vel[t][d] = 0.0; wm.omega[t][d] = 0.0; }
            wm.omega[t][2] = dtheta_dt;
            const double Rz[3][3] = {{cosD, -sinD, 0.0}, {sinD, cosD, 0.0}, {0.0, 0.0, 1.0}};
            for (int d = 0; d < 3; ++d)
                for (int e = 0; e < 3; ++e) wm.R[t][d][e] = Rz[d][e];
        }
    } else if (wm.active)
        for (int t = 0; t < kTypeCount; ++t)
            for (int d = 0; d < 3; ++d) {
                wm.center[t][d] = c->wall_center[t][d];
                wm.vel[t][d] = c->p.wall_velocity[t][d];
                wm.omega[t][d] = c->p.wall_omega[t][d];
                for (int e = 0; e < 3; ++e) wm.R[t][d][e] = c->c.wall_rotation[t][d][e];
            }
    const DecideArgs da = decide_args(c); // (before the wall centres advance: walls_move looks at this step's Time)
    const bool defer = defer_solids(c);
    LAUNCH(c, k_prestep, nb, kBlock, ctl, c->S, c->sol, c->grid, wm, motion ? 1 : 0, c->ancx, c->ancy, c->ancz, (const int *)nullptr, 0,
           defer ? 1 : 0);
    if (defer) { // the solids' share, once their sub-steps have finished
        int rc = join_solids(c);
        if (rc) return rc;
        LAUNCH(c, k_prestep, nblk(c->ns), kBlock, ctl, c->S, c->sol, c->grid, wm, motion ? 1 : 0, c->ancx, c->ancy, c->ancz,
               (const int *)c->sol.slot, c->ns, 0);
        if (early_pass1 && c->early_pass1) {
            CK(cudaEventRecord(c->ev_u_ready, c->stream));
            CK(cudaStreamWaitEvent(c->side, c->ev_u_ready, 0));
            if ((rc = run_solid_substeps(c, c->side, 1))) return rc;
            c->early_done = true;
        }
    }
    if (motion) // :3066-3070 (host mirror of the wall centres)
        for (int t = 4; t < kTypeCount; ++t)
            for (int d = 0; d < 3; ++d) c->wall_center[t][d] += c->p.wall_velocity[t][d] * c->p.dt;
    if (c->tracing) LAUNCH(c, k_mark, 1, 1, ctl, 9); // (pre-step done, incl. the wait for the previous step's sub-steps)
    // rebuild or reuse: every rank votes, the OR decides (one context: its own vote)
    if (c->slab) {
        LAUNCH(c, k_vote, 1, 32, ctl, c->epoch, c->peers, da, (const int *)c->pl.flags);
        LAUNCH(c, k_wait_decide, 1, 32, ctl, c->epoch, c->mine.vote, c->nranks, da, c->pl.flags);
    } else {
        LAUNCH(c, k_need, 1, 1, ctl, da, (const int *)c->pl.flags);
        LAUNCH(c, k_decide, 1, 1, ctl, da, c->pl.flags);
    }
    if (c->slab) { // a step that reuses its list sends the halo particles' new state first: the rebuild-only kernels below return at once
        k_halo_resend<<<dim3(kPushBlocks, 2), kPushThreads, 0, c->stream>>>(ctl, c->epoch, kPushHaloL, c->S, c->haloSlot[0], c->haloSlot[1],
                                                                          push_pair(c, true));
        ++c->launches;
    }
    SlabSend snd{};
    snd.buf[0] = c->stage_mig[0]; snd.buf[1] = c->stage_mig[1]; snd.capacity = c->msg_cap;
    LAUNCH(c, k_count, small_grid(c->nmax), kBlock, ctl, c->S, c->grid, c->cellCount, c->slot, snd);
    if (c->slab) {
        int rc;
        const dim3 mb2(small_grid(c->msg_cap), 2);
        const double W0 = c->grid.W[0];
        // (1) migration (rebuild steps only)
        if ((rc = exchange_particles(c, false))) return rc;
        SidePair mg{};
        mg.buf[0] = c->mine.mig[0]; mg.buf[1] = c->mine.mig[1];
        k_unpack_particles<<<mb2, kBlock, 0, c->stream>>>(ctl, mg, c->mine.cnt_mig, c->msg_cap, c->cap, 0, c->S, c->grid, c->cellCount, c->slot,
                                                          c->ancx, c->ancy, c->ancz);
        ++c->launches;
        LAUNCH(c, k_advance_n, 1, 1, ctl, c->mine.cnt_mig, c->msg_cap, c->cap, 0);
        // (2) halo: rebuild steps pack the particles within one halo width of the faces, reuse steps re-send the same ones
        SlabSend hs{};
        hs.buf[0] = c->stage_halo[0]; hs.buf[1] = c->stage_halo[1]; hs.capacity = c->msg_cap;
        LAUNCH(c, k_halo_pack, small_grid(c->nmax), kBlock, ctl, c->S, c->grid, hs, c->haloSrc[0], c->haloSrc[1]);
        if ((rc = exchange_particles(c, true))) return rc;
        // halo copies that crossed the periodic seam are shifted by the box width so that separations in
        // x need no wrap inside a slab; migrants were already wrapped into the box by the sender
        SidePair hl{};
        hl.buf[0] = c->mine.halo[0]; hl.buf[1] = c->mine.halo[1];
        hl.xshift[0] = c->rank == 0 ? -W0 : 0.0; hl.xshift[1] = c->rank == c->nranks - 1 ? W0 : 0.0;
        k_unpack_refresh<<<mb2, kBlock, 0, c->stream>>>(ctl, hl, c->S, c->ghostSlot[0], c->ghostSlot[1]);
        k_unpack_particles<<<mb2, kBlock, 0, c->stream>>>(ctl, hl, c->mine.cnt_halo, c->msg_cap, c->cap, kGhost, c->S, c->grid, c->cellCount, c->slot,
                                                          c->ancx, c->ancy, c->ancz);
        c->launches += 2;
        LAUNCH(c, k_advance_n, 1, 1, ctl, c->mine.cnt_halo, c->msg_cap, c->cap, 1);
    }
    if (c->tracing) LAUNCH(c, k_mark, 1, 1, ctl, 10); // (migration + halo exchanged)
    const int nc = c->grid.ncells + 2; // + parked + dead buckets
    LAUNCH(c, k_scan_reduce, c->scan_blocks, kScanThreads, ctl, c->cellCount, nc, c->blockSums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, ctl, c->blockSums, c->scan_blocks);
    LAUNCH(c, k_scan_apply, c->scan_blocks, kScanThreads, ctl, c->cellCount, nc, c->blockSums, c->cellStart, 1);
    LAUNCH(c, k_scatter_index, small_grid(c->nmax), kBlock, ctl, c->S.key, c->slot, c->cellStart, c->tmpIdx);
    LAUNCH(c, k_permute, nb, kBlock, ctl, c->S, c->T, c->cellStart, c->tmpIdx, c->grid, c->where, c->sol.slot, c->sol.sb, c->ancx, c->ancy,
           c->ancz);
    LAUNCH(c, k_set_n, 1, 1, ctl, c->cellStart, c->grid.ncells);
    std::swap(c->S, c->T);
    c->bx = c->S.x; c->by = c->S.y; c->bz = c->S.z;
    if (c->slab) {
        LAUNCH(c, k_slab_slots, small_grid(c->msg_cap), kBlock, ctl, c->where, c->haloSrc[0], c->haloSrc[1], c->haloSlot[0], c->haloSlot[1],
               c->ghostSlot[0], c->ghostSlot[1]);
        if (c->ns > 0) LAUNCH(c, k_solid_owned_list, small_grid(c->ns), kBlock, ctl, c->S, c->sol, c->own_sol);
    }
    CK(cudaGetLastError());
    return MPHX_OK;
}

static int run_pass1(Ctx *c, bool timed = false)
{
    Ctl *ctl = c->ctl;
    const int nmax = c->nmax;
    const int batch = c->p.dim == 3 ? c->sweep_batch : c->grid.nsten;
    if (c->pl.nbr) { // K5a: the candidate list (rebuild steps only: the kernel returns at once otherwise)
        const int npairs = (nmax + 1) / 2;
        // 32-bit offsets must hold a parked top plus one stride ((L + 1) * cap + i + cap) without wrapping
        const bool big = ((size_t)c->pl.L + 2) * (size_t)c->pl.cap >= 0xffffffffull;
        if (c->filter2 || big) {
#define F2(D, B) LAUNCH(c, (k_filter2<D, B>), nblk(npairs, kSweepThreads), kSweepThreads, ctl, c->S, c->cellStart, c->grid, c->pl)
            if (c->p.dim == 3) { if (big) F2(3, true); else F2(3, false); }
            else               { if (big) F2(2, true); else F2(2, false); }
#undef F2
        } else {
            if (c->p.dim == 3) LAUNCH(c, k_filter<3>, nblk(nmax, kSweepThreads), kSweepThreads, ctl, c->S, c->cellStart, c->grid, c->pl);
            else               LAUNCH(c, k_filter<2>, nblk(nmax, kSweepThreads), kSweepThreads, ctl, c->S, c->cellStart, c->grid, c->pl);
        }
    }
    PairList pl1 = c->pl; // (pass 1's view of the list: with the staged path on, its particles are masked out of the list kernel)
    if (c->brick && c->lnbr) {
        LAUNCH(c, k_brick_count, nblk(c->nbricks), kBlock, ctl, c->cellStart, c->grid, c->nbricks, c->brick_nown);
        LAUNCH(c, k_scan_reduce, c->brick_scan_blocks, kScanThreads, ctl, c->brick_nown, c->nbricks, c->brick_sums);
        LAUNCH(c, k_scan_top, 1, kScanThreads, ctl, c->brick_sums, c->brick_scan_blocks);
        LAUNCH(c, k_scan_apply, c->brick_scan_blocks, kScanThreads, ctl, c->brick_nown, c->nbricks, c->brick_sums, c->brick_base, 0);
        LAUNCH(c, k_brick_localize, c->nbricks, 256, ctl, c->S, c->cellStart, c->grid, c->pl, c->lnbr, c->brick_base, c->brick_ok, c->in_brick);
        pl1.skip = c->in_brick;
    }
    if (timed) timer_mark(c);
    if (c->brick && c->lnbr) {
        k_brick_pass1<3><<<c->nbricks, kBrickThreads, kBrickCap * (sizeof(Rec) + sizeof(double2)), c->stream>>>(
            ctl, c->S, c->cellStart, c->grid, c->phys, c->pl, c->lnbr, c->brick_base, c->brick_ok, c->P, c->volStrain, c->divP);
        ++c->launches;
    }
    // fused-sweep fall-backs: a small persistent grid when they only have to look at the overflow flag
    const int vblocks = nblk(nmax, kSweepThreads);
#define SWEEP_GRID(LIST) ((LIST) || !c->pl.nbr ? vblocks : std::min(vblocks, 4 * 148))
#define P1(D, ST, LIST) LAUNCH(c, (k_pass1_v3<D, ST, LIST>), SWEEP_GRID(LIST), kSweepThreads, ctl, c->S, c->cellStart, c->grid, c->phys, \
                         batch, c->P, c->volStrain, c->divP, c->densA, c->gcx, c->gcy, c->gcz, c->PA, pl1)
#define P1D(ST, LIST) do { if (c->p.dim == 3) P1(3, ST, LIST); else P1(2, ST, LIST); } while (0)
    if (c->pl.nbr) { // list traversal, then the fused sweep for particles whose list overflowed (normally none)
        if (c->surface_tension) P1D(true, true); else P1D(false, true);
    }
    if (c->surface_tension) P1D(true, false); else P1D(false, false);
#undef P1D
#undef P1
    CK(cudaGetLastError());
    return MPHX_OK;
}

// slab mode, between the passes: PressureP of the halo copies (ring neighbours) and of the replicated solids (all ranks)
static int exchange_pressure(Ctx *c)
{
    Ctl *ctl = c->ctl;
    Mailbox &L = c->peers.left, &R = c->peers.right;
    PassOneFields pf{};
    PassOneTargets pt{};
    pf.a[0] = c->P; pt.a[0] = c->P; pf.count = pt.count = 1;
    if (c->surface_tension) { // PressureA and GravityCenter of the halo copies feed pass 2's surface-tension terms
        double *st[4] = {c->PA, c->gcx, c->gcy, c->gcz};
        for (int p = 0; p < 4; ++p) { pf.a[1 + p] = st[p]; pt.a[1 + p] = st[p]; }
        pf.count = pt.count = 5;
    }
    ScalarPush sp{};
    sp.dst[0] = L.p[1]; sp.dst[1] = R.p[0]; sp.dst_flag[0] = L.fp + 1; sp.dst_flag[1] = R.fp + 0;
    k_push_scalar<<<dim3(kPushBlocks, 2), kPushThreads, 0, c->stream>>>(ctl, c->epoch, kPushPL, c->haloSlot[0], c->haloSlot[1], pf, c->msg_cap, sp);
    ++c->launches;
    if (c->ns > 0) LAUNCH(c, k_solid_publish_P, kPushBlocks, kPushThreads, ctl, c->epoch, kPushSolP, c->S, c->sol, c->own_sol, c->P, c->peers);
    // one wait for the neighbours' PressureP and every rank's share of the replicated solids' PressureP
    LAUNCH(c, k_wait<0>, 1, 32, ctl, c->epoch, c->mine.fp, 2, kWaitP, 0, c->ns > 0 ? c->mine.fsolP : (const unsigned long long *)nullptr,
           c->ns > 0 ? c->nranks : 0);
    k_unpack_scalar<<<dim3(small_grid(c->msg_cap), 2), kBlock, 0, c->stream>>>(ctl, c->ghostSlot[0], c->ghostSlot[1], c->mine.p[0], c->mine.p[1],
                                                                             c->msg_cap, pt, c->S.rb);
    ++c->launches;
    if (c->ns > 0)
        LAUNCH(c, k_solid_spread_P, small_grid(c->ns), kBlock, ctl, c->S, c->grid, c->sol, c->mine.solP, c->P,
               c->surface_tension ? c->PA : (double *)nullptr, c->gcx, c->gcy, c->gcz, c->phys);
    CK(cudaGetLastError());
    return MPHX_OK;
}

// the solids' share of pass 2 first (a few blocks), so that the sub-steps can start on the second stream
static bool solid_split(const Ctx *c) { return defer_solids(c); }

static int run_pass2(Ctx *c)
{
    Ctl *ctl = c->ctl;
    const int nmax = c->nmax;
    const int batch = c->p.dim == 3 ? c->sweep_batch : c->grid.nsten;
    const int vblocks = nblk(nmax, kSweepThreads);
#define P2(D, ST, LIST, GRID, SUB) LAUNCH(c, (k_pass2_v3<D, ST, LIST>), GRID, kSweepThreads, ctl, c->S, c->cellStart, c->grid, \
                         c->phys, batch, c->P, c->PA, c->gcx, c->gcy, c->gcz, c->T.x, c->T.y, c->T.z, c->T.vx, c->T.vy, c->T.vz, c->fx, \
                         c->fy, c->fz, c->ax, c->ay, c->az, c->sol, c->pl, SUB)
#define P2D(ST, LIST, GRID, SUB) do { if (c->p.dim == 3) P2(3, ST, LIST, GRID, SUB); else P2(2, ST, LIST, GRID, SUB); } while (0)
#define P2ALL(GRIDL, SUB) do { \
        if (c->pl.nbr) { /* list traversal, then the sweep variant for particles whose list overflowed (normally none) */ \
            if (c->surface_tension) P2D(true, true, GRIDL, SUB); else P2D(false, true, GRIDL, SUB); \
        } \
        if (c->surface_tension) P2D(true, false, SWEEP_GRID(false), SUB); else P2D(false, false, SWEEP_GRID(false), SUB); \
    } while (0)
    const bool split = solid_split(c);
    if (c->ns > 0 && (split || c->slab)) {
        const Subset solids{c->sol.slot, c->ns, 0};
        P2ALL(nblk(c->ns, kSweepThreads), solids);
        // slab mode: the owners' coupled velocities go to every rank's mailbox
        if (c->slab) LAUNCH(c, k_solid_publish_V, kPushBlocks, kPushThreads, ctl, c->epoch, kPushSolV, c->S, c->sol, c->own_sol, c->peers);
        if (split) CK(cudaEventRecord(c->ev_solid_ready, c->stream));
        if (c->tracing) LAUNCH(c, k_mark, 1, 1, ctl, 5);
        const Subset rest{nullptr, 0, 1};
        P2ALL(vblocks, rest);
    } else {
        const Subset all{nullptr, 0, 0};
        P2ALL(vblocks, all);
    }
#undef P2ALL
#undef P2D
#undef P2
#undef SWEEP_GRID
    // S keeps type/id/key of this step's order and takes the integrated x,v; T keeps the pre-step
    // positions (the ones the reference's NeighborCount refers to).
    std::swap(c->S.x, c->T.x); std::swap(c->S.y, c->T.y); std::swap(c->S.z, c->T.z);
    std::swap(c->S.vx, c->T.vx); std::swap(c->S.vy, c->T.vy); std::swap(c->S.vz, c->T.vz);
    c->bx = c->T.x; c->by = c->T.y; c->bz = c->T.z;
    CK(cudaGetLastError());
    return MPHX_OK;
}

// part 0: the whole of the step's sub-steps; 1: only pass 1 of the first sub-step (it reads the displacements the pre-step
// left, not the coupled velocities: on the second stream it starts right after the solids' pre-step, beside the bucket stage and
// the fluid passes); 2: the rest (the owners' coupled velocities arrive first)
static int run_solid_substeps(Ctx *c, cudaStream_t strm, int part = 0)
{
    if (c->ns <= 0) return MPHX_OK;
    const int substeps = (int)(c->p.dt / c->p.elastic_dt + 0.5); // :653
    const mphx_constants &k = c->c;
    const int ns = c->ns;
    // Q1: the second position update is the `#else` tail of the Rolling2 block (:2070-2079): every variant but Rolling2
    const int dbl = ((c->p.ref_compat & MPHX_COMPAT_DOUBLE_UPDATE) && c->p.clamp_module != MPHX_MODULE_ROLLING2) ? 1 : 0;
    // slab mode: the owners' coupled velocities arrive, then either every rank runs identical sub-steps of all solids, or
    // (split sub-steps, the default) every rank advances its share [s_lo, s_hi) and the ranks trade P and u after every
    // pass: 2 * substeps phases, counted by epoch * kSubPhases + phase in every rank's fsub flags
    const bool ring = c->slab && c->split_substeps;
    SolidRing rg = c->ring;
    if (!ring) rg.nranks = 0;
    if (ring && 2 * substeps + 1 >= kSubPhases) { set_last_error("too many solid sub-steps per step for the split exchange"); return MPHX_ERR_UNSUPPORTED; }
    const unsigned long long seq0 = c->epoch * (unsigned long long)kSubPhases;
    const int cnt = c->s_hi - c->s_lo;
    const bool deep = ring || (c->sol.packed && std::getenv("MPHX_DEBUG_DEEP") != nullptr); // (the variable: developer timing of the deep variant)
    const int threads = deep ? kRingBlock : kBlock;
    const int grid = ring ? std::max(nblk(cnt, threads), 1) : nblk(cnt, threads); // (an empty share still posts its phase counters)
    // Measured on B200 (tools/substep_bench.py, 113k solids, ms per step of 5 sub-steps, kernels alone):
    //   share of the solids      1/1     1/2     1/4     1/8
    //   one thread per solid     0.608   0.404   0.276   0.264     <- a latency chain of ~80 / ~160 dependent gathers per thread
    //   ... 8 gathers in flight  0.681   0.416   0.195   0.162     <- DEEP: what a rank's share uses (ring)
    //   team of 16 lanes         1.619   0.839   0.391   0.233     <- 5x the instructions and L1 wavefronts (a team's gathers do not
    //                                                                  coalesce across lanes): kept for MPHX_SOLID_TEAM=2 only
    const bool team = c->team_ok && c->sol.packed && c->solid_team_always;
    const int tgrid = std::max(nblk(cnt, kTeamBlock / kTeam), 1);
#define PASS1(D, PK, RG, DP) LAUNCH_ON(c, strm, (k_solid_pass1<D, PK, RG, DP>), grid, threads, c->ctl, c->sol, c->s_lo, c->s_hi, rg)
#define PASS2(D, PK, RG, DP) LAUNCH_ON(c, strm, (k_solid_pass2<D, PK, RG, DP>), grid, threads, c->ctl, c->sol, c->s_lo, c->s_hi, k.domain_width[0], \
                                       k.domain_width[1], k.domain_width[2], c->p.elastic_dt, c->p.clamp_module, dbl, c->d_inv_density, rg)
#define TEAM1(D, RG) do { k_solid_pass1_team<D, RG><<<tgrid, kTeamBlock, 0, strm>>>(c->ctl, c->sol, c->s_lo, c->s_hi, rg); ++c->launches; } while (0)
#define TEAM2(D, RG) do { k_solid_pass2_team<D, RG><<<tgrid, kTeamBlock, 0, strm>>>(c->ctl, c->sol, c->s_lo, c->s_hi, k.domain_width[0],            \
                              k.domain_width[1], k.domain_width[2], c->p.elastic_dt, c->p.clamp_module, dbl, c->d_inv_density, rg); ++c->launches; } while (0)
#define PHASE_D(D, WHICH)                                                                                                              \
    do {                                                                                                                                \
        if (team) { if (ring) TEAM##WHICH(D, true); else TEAM##WHICH(D, false); }                                                       \
        else if (c->sol.packed) { if (ring) PASS##WHICH(D, true, true, true); else if (deep) PASS##WHICH(D, true, false, true);         \
                                  else PASS##WHICH(D, true, false, false); }                                                            \
        else { if (ring) PASS##WHICH(D, false, true, true); else PASS##WHICH(D, false, false, false); }                                 \
    } while (0)
    const int first = part == 2 ? 1 : 0, last = part == 1 ? 1 : 2 * substeps;
    for (int ph = first; ph < last; ++ph) {
        if (c->slab && ph == (part == 0 ? 0 : 1)) { // the owners' coupled velocities (pass 2 needs them, pass 1 does not)
            LAUNCH_ON(c, strm, k_wait<0>, 1, 32, c->ctl, c->epoch, c->mine.fsolV, c->nranks, kWaitSolV);
            LAUNCH_ON(c, strm, k_solid_apply_update, nblk(ns), kBlock, c->sol, c->mine.solV, c->s_lo, c->s_hi, c->p.clamp_module);
        }
        rg.seq = seq0 + 1 + ph;
        // every rank's previous phase: waited for by a one-warp kernel after each phase (default), or inside the kernel that
        // needs it (MPHX_RING_INKERNEL_WAIT=1: one launch per phase; measured equal at 8 GPUs, 2 % slower at 2, where a share
        // is ~1800 one-warp blocks that hold their registers while they spin)
        rg.wait_seq = (ring && c->ring_inkernel_wait && ph > 0) ? seq0 + ph : 0;
        rg.last = ph == 2 * substeps - 1 ? 1 : 0;
        if (ph % 2 == 0) { if (c->p.dim == 3) PHASE_D(3, 1); else PHASE_D(2, 1); }
        else             { if (c->p.dim == 3) PHASE_D(3, 2); else PHASE_D(2, 2); }
        // (after the step's last phase in any case: everybody's final state has arrived before anything reads the solids)
        if (ring && (!c->ring_inkernel_wait || ph == 2 * substeps - 1))
            LAUNCH_ON(c, strm, k_wait<0>, 1, 32, c->ctl, rg.seq, c->mine.fsub, c->nranks, kWaitSub);
    }
#undef PHASE_D
#undef TEAM2
#undef TEAM1
#undef PASS2
#undef PASS1
    CK(cudaGetLastError());
    return MPHX_OK;
}

static cudaEvent_t timer_event(Ctx *c)
{
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
static void timer_mark(Ctx *c)
{
    if (!c->timing) return;
    cudaEvent_t e = timer_event(c);
    cudaEventRecord(e, c->stream);
    c->ev.push_back(e);
}
static void timer_resolve(Ctx *c)
{
    if (c->ev.empty() && c->side_marks.empty()) return;
    cudaStreamSynchronize(c->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    // events come in groups of 6 per step: start, after rebuild, after the filter, after pass 1, after
    // pass 2, after the solid sub-steps (when those run on the context's stream)
    for (size_t i = 0; i + 5 < c->ev.size(); i += 6)
        for (int k = 0; k < 5; ++k) {
            float a = 0;
            cudaEventElapsedTime(&a, c->ev[i + k], c->ev[i + k + 1]);
            c->ms[k] += a;
        }
    // sub-steps on the second stream: their own event pairs (the reference's "explicit calculation" timer
    // includes them, src/main.cpp:669)
    for (size_t i = 0; i + 1 < c->side_marks.size(); i += 2) {
        float a = 0;
        cudaEventElapsedTime(&a, c->side_marks[i], c->side_marks[i + 1]);
        c->ms[4] += a;
    }
    for (cudaEvent_t e : c->ev) c->ev_pool.push_back(e);
    for (cudaEvent_t e : c->side_marks) c->ev_pool.push_back(e);
    c->ev.clear();
    c->side_marks.clear();
}

#define TRACE_ON(strm, code) do { if (c->tracing) LAUNCH_ON(c, strm, k_mark, 1, 1, c->ctl, code); } while (0)
static int apply_recut(Ctx *c);
static int one_step(Ctx *c, bool fluid_only)
{
    int rc;
    if (c->slab && !c->connected) { set_last_error("slab context is not connected to its peers (mphx_slab_connect)"); return MPHX_ERR_INVALID; }
    if (c->pending_lo >= 0 && (rc = apply_recut(c))) return rc;
    timer_mark(c);
    TRACE_ON(c->stream, 1);
    c->early_done = false;
    if ((rc = stage_build(c, true, !fluid_only))) return rc; // calculateWall, PeriodicBoundary, resets, calculateNeighbor
    timer_mark(c);
    TRACE_ON(c->stream, 2);
    if ((rc = run_pass1(c, true))) return rc;       // filter; DensityA..DivergenceP, coefficients, PressureP/A
    TRACE_ON(c->stream, 3);
    if (c->slab && (rc = exchange_pressure(c))) return rc;
    timer_mark(c);
    TRACE_ON(c->stream, 4);
    if ((rc = run_pass2(c))) return rc;             // force sums, gravity, interface, acceleration, convection
    timer_mark(c);
    TRACE_ON(c->stream, 6);
    if (!fluid_only) {
        if (solid_split(c)) { // sub-steps on the second stream: joined by the next pre-step (or by any reader)
            CK(cudaStreamWaitEvent(c->side, c->ev_solid_ready, 0));
            TRACE_ON(c->side, 7);
            if (c->timing) { cudaEvent_t e = timer_event(c); cudaEventRecord(e, c->side); c->side_marks.push_back(e); }
            if ((rc = run_solid_substeps(c, c->side, c->early_done ? 2 : 0))) return rc;
            TRACE_ON(c->side, 8);
            if (c->timing) { cudaEvent_t e = timer_event(c); cudaEventRecord(e, c->side); c->side_marks.push_back(e); }
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

// In-place re-balancing of the slabs: this slab's owned columns become [pending_lo, pending_hi).  A face moves by
// at most one halo width per re-cut, so the particles that change owner are exactly what the ordinary MIGRATION of the next
// (forced) rebuild hands to the ring neighbour -- no host gather, no new exchange path: the grid descriptor changes, the
// bucket arrays grow if they must, and the step that follows rebuilds.  Applied at the start of a step so that everything
// read between the request and the step (downloads, ownership masks) still refers to the grid the keys were made on.
static int apply_recut(Ctx *c)
{
    const int lo = c->pending_lo, hi = c->pending_hi;
    c->pending_lo = c->pending_hi = -1;
    int rc;
    if ((rc = join_solids(c))) return rc;
    GridDesc &g = c->grid;
    const int R = g.range;
    const long long nc = (long long)((hi - lo) + 2 * R) * g.ny * g.nz;
    if (nc > 2000000000LL) return MPHX_ERR_UNSUPPORTED;
    if ((size_t)nc > c->cells_cap) { // (rare: the old arrays stay with the context until it is destroyed)
        CK(cudaStreamSynchronize(c->stream));
        const size_t cap = (size_t)nc + (size_t)nc / 8;
        int *cc = nullptr, *cs = nullptr, *bs = nullptr;
        if (c->alloc(&cc, cap + 2) || c->alloc(&cs, cap + 3) || c->alloc(&bs, (cap + 2 + kScanChunk - 1) / kScanChunk + 1)) return MPHX_ERR_NOMEM;
        CK(cudaMemsetAsync(cc, 0, sizeof(int) * (cap + 2), c->stream)); // (the counters are all-zero outside a rebuild)
        c->cellCount = cc; c->cellStart = cs; c->blockSums = bs; c->cells_cap = cap;
    }
    c->col_lo = lo; c->col_hi = hi;
    g.nx = (hi - lo) + 2 * R;
    g.xoff = lo - R;
    g.mn[0] = g.mn0g + (double)g.xoff * g.cellw;
    g.ncells = (int)nc;
    c->scan_blocks = (int)((nc + 2 + kScanChunk - 1) / kScanChunk);
    return request_rebuild(c);
}

// ---- exact neighbour sets (debug / VTK NeighborCount / the initial structure lists) -----------------------
// A private bucket structure over `n` positions (x, y, z; type, id) on `grid`, then the reference's
// bit-exact predicate.  Independent of the stepping state (which may be several steps into a reused list).
static int exact_lists(Ctx *c, int n, const double *x, const double *y, const double *z, const int *type, const int *id, GridDesc grid,
                       bool structure_only, bool xy_only, int row_base, int nrows, std::vector<long long> &offsets, std::vector<int> *ids_out)
{
    Scratch tmp;
    double *sx, *sy, *sz;
    int *stype, *sid, *skey, *key, *slot, *cellCount, *cellStart, *blockSums, *d_counts;
    const size_t nn = (size_t)std::max(n, 1), nc = (size_t)grid.ncells + 2;
    const int sb_blocks = (int)((nc + kScanChunk - 1) / kScanChunk);
    int e = 0;
    e |= tmp.get(&sx, nn); e |= tmp.get(&sy, nn); e |= tmp.get(&sz, nn);
    e |= tmp.get(&stype, nn); e |= tmp.get(&sid, nn); e |= tmp.get(&skey, nn); e |= tmp.get(&key, nn); e |= tmp.get(&slot, nn);
    e |= tmp.get(&cellCount, nc); e |= tmp.get(&cellStart, nc + 1); e |= tmp.get(&blockSums, (size_t)sb_blocks + 1);
    e |= tmp.get(&d_counts, (size_t)std::max(nrows, 1));
    if (e) return MPHX_ERR_NOMEM;
    CK(cudaMemsetAsync(cellCount, 0, sizeof(int) * nc, c->stream));
    CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)std::max(nrows, 1), c->stream));
    LAUNCH(c, k_dbg_keycount, nblk(n), kBlock, n, x, y, z, grid, key, cellCount, slot);
    LAUNCH(c, k_scan_reduce, sb_blocks, kScanThreads, (const Ctl *)nullptr, cellCount, (int)nc, blockSums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, (const Ctl *)nullptr, blockSums, sb_blocks);
    LAUNCH(c, k_scan_apply, sb_blocks, kScanThreads, (const Ctl *)nullptr, cellCount, (int)nc, blockSums, cellStart, 0);
    LAUNCH(c, k_dbg_gather, nblk(n), kBlock, n, x, y, z, type, id, key, slot, cellStart, sx, sy, sz, stype, sid, skey);
    const double cut = c->c.max_radius + 0.1 * c->p.particle_spacing;
    const double cutoff2 = cut * cut; // (MaxRadius+MARGIN)*(MaxRadius+MARGIN) :1765
    if (grid.dim == 3)
        LAUNCH(c, (k_neighbors_exact<3, 0>), nblk(n), kBlock, n, sx, sy, sz, stype, sid, skey, cellStart, grid, cutoff2, structure_only ? 1 : 0,
               xy_only ? 1 : 0, row_base, d_counts, (const long long *)nullptr, (int *)nullptr);
    else
        LAUNCH(c, (k_neighbors_exact<2, 0>), nblk(n), kBlock, n, sx, sy, sz, stype, sid, skey, cellStart, grid, cutoff2, structure_only ? 1 : 0,
               xy_only ? 1 : 0, row_base, d_counts, (const long long *)nullptr, (int *)nullptr);
    std::vector<int> counts((size_t)std::max(nrows, 1));
    CK(cudaMemcpyAsync(counts.data(), d_counts, sizeof(int) * (size_t)std::max(nrows, 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    offsets.assign((size_t)nrows + 1, 0);
    for (int r = 0; r < nrows; ++r) offsets[r + 1] = offsets[r] + counts[r];
    const long long total = offsets[nrows];
    if (!ids_out) { CK(cudaGetLastError()); return MPHX_OK; }
    int *d_ids;
    long long *d_off;
    if (tmp.get(&d_ids, (size_t)std::max<long long>(total, 1)) || tmp.get(&d_off, (size_t)nrows + 1)) return MPHX_ERR_NOMEM;
    CK(cudaMemcpyAsync(d_off, offsets.data(), sizeof(long long) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, c->stream));
    if (grid.dim == 3)
        LAUNCH(c, (k_neighbors_exact<3, 1>), nblk(n), kBlock, n, sx, sy, sz, stype, sid, skey, cellStart, grid, cutoff2, structure_only ? 1 : 0,
               xy_only ? 1 : 0, row_base, (int *)nullptr, d_off, d_ids);
    else
        LAUNCH(c, (k_neighbors_exact<2, 1>), nblk(n), kBlock, n, sx, sy, sz, stype, sid, skey, cellStart, grid, cutoff2, structure_only ? 1 : 0,
               xy_only ? 1 : 0, row_base, (int *)nullptr, d_off, d_ids);
    ids_out->assign((size_t)std::max<long long>(total, 1), 0);
    CK(cudaMemcpyAsync(ids_out->data(), d_ids, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    ids_out->resize((size_t)total);
    return MPHX_OK;
}
// neighbour sets of the particles this context holds, over the positions of the last pass 1
static int exact_lists_current(Ctx *c, std::vector<long long> &offsets, std::vector<int> *ids_out)
{
    int n = 0, rc;
    if ((rc = join_solids(c))) return rc;
    if ((rc = read_n(c, &n))) return rc;
    return exact_lists(c, n, c->bx, c->by, c->bz, c->S.type, c->S.id, c->grid, false, false, 0, c->n_global, offsets, ids_out);
}

// calculateVirialStressAtParticle (:3077-3318) on the state after the last step, over the reference's lists of that
// step (pre-step positions c->bx, bit-exact predicate); results in ORIGINAL particle order on the device
static int compute_virial(Ctx *c, double *d_out9, double *d_outp)
{
    if (c->slab) { set_last_error("the virial stress diagnostic needs the post-step state of the halo: single context only"); return MPHX_ERR_UNSUPPORTED; }
    int n = 0, rc;
    if ((rc = join_solids(c))) return rc;
    if ((rc = read_n(c, &n))) return rc;
    Scratch tmp;
    double *sx, *sy, *sz, *cur[6];
    int *stype, *sslot, *skey, *key, *slot, *iota, *cellCount, *cellStart, *blockSums;
    const GridDesc &grid = c->grid;
    const size_t nn = (size_t)std::max(n, 1), nc = (size_t)grid.ncells + 2;
    const int sb_blocks = (int)((nc + kScanChunk - 1) / kScanChunk);
    int e = 0;
    e |= tmp.get(&sx, nn); e |= tmp.get(&sy, nn); e |= tmp.get(&sz, nn);
    for (double *&q : cur) e |= tmp.get(&q, nn);
    e |= tmp.get(&stype, nn); e |= tmp.get(&sslot, nn); e |= tmp.get(&skey, nn); e |= tmp.get(&key, nn); e |= tmp.get(&slot, nn); e |= tmp.get(&iota, nn);
    e |= tmp.get(&cellCount, nc); e |= tmp.get(&cellStart, nc + 1); e |= tmp.get(&blockSums, (size_t)sb_blocks + 1);
    if (e) return MPHX_ERR_NOMEM;
    CK(cudaMemsetAsync(cellCount, 0, sizeof(int) * nc, c->stream));
    CK(cudaMemsetAsync(d_out9, 0, sizeof(double) * 9 * (size_t)c->n_global, c->stream));
    CK(cudaMemsetAsync(d_outp, 0, sizeof(double) * (size_t)c->n_global, c->stream));
    LAUNCH(c, k_iota, nblk(n), kBlock, n, iota);
    LAUNCH(c, k_dbg_keycount, nblk(n), kBlock, n, c->bx, c->by, c->bz, grid, key, cellCount, slot);
    LAUNCH(c, k_scan_reduce, sb_blocks, kScanThreads, (const Ctl *)nullptr, cellCount, (int)nc, blockSums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, (const Ctl *)nullptr, blockSums, sb_blocks);
    LAUNCH(c, k_scan_apply, sb_blocks, kScanThreads, (const Ctl *)nullptr, cellCount, (int)nc, blockSums, cellStart, 0);
    LAUNCH(c, k_dbg_gather, nblk(n), kBlock, n, c->bx, c->by, c->bz, c->S.type, iota, key, slot, cellStart, sx, sy, sz, stype, sslot, skey);
    LAUNCH(c, k_current_state, nblk(n), kBlock, n, c->S, c->sol, cur[0], cur[1], cur[2], cur[3], cur[4], cur[5]);
    VirialIn in{};
    in.sx = sx; in.sy = sy; in.sz = sz; in.sslot = sslot; in.skey = skey; in.cellStart = cellStart;
    in.cx = cur[0]; in.cy = cur[1]; in.cz = cur[2]; in.cvx = cur[3]; in.cvy = cur[4]; in.cvz = cur[5];
    in.type = c->S.type; in.id = c->S.id; in.P = c->P; in.PA = c->PA; in.gcx = c->gcx; in.gcy = c->gcy; in.gcz = c->gcz;
    in.out9 = d_out9; in.outp = d_outp;
    const double cut = c->c.max_radius + 0.1 * c->p.particle_spacing;
    if (grid.dim == 3) LAUNCH(c, k_virial<3>, nblk(n), kBlock, n, in, grid, c->phys, cut * cut, c->surface_tension ? 1 : 0);
    else               LAUNCH(c, k_virial<2>, nblk(n), kBlock, n, in, grid, c->phys, cut * cut, c->surface_tension ? 1 : 0);
    CK(cudaStreamSynchronize(c->stream)); // (the scratch buffers are released on return)
    CK(cudaGetLastError());
    return MPHX_OK;
}

// ---- initial structure lists (calculateInitialNeighbor :1497-1644) + Lame + Normalizer ----------
static int init_solid(Ctx *c)
{
    if (c->ns <= 0) return MPHX_OK;
    const int ns = c->ns;
    int rc;
    // calculateInitialNeighbor (:1497-1644): buckets over InitialPosition, structure particles only, on the
    // GLOBAL periodic grid, so the lists are complete on every slab of a multi-GPU run.
    GridDesc g = c->grid;
    {
        const mphx_constants &k = c->c;
        g.slab = 0; g.nx = k.cell_count[0]; g.nxg = g.nx; g.xoff = 0; g.mn[0] = c->p.domain_min[0]; g.mn0g = g.mn[0];
        g.ncells = k.cell_counts;
    }
    std::vector<long long> off;
    std::vector<int> ids;
    {
        Scratch tmp;
        int *sid;
        if (tmp.get(&sid, (size_t)ns)) return MPHX_ERR_NOMEM;
        std::vector<int> h((size_t)ns);
        for (int s = 0; s < ns; ++s) h[s] = c->sol.sb + s;
        CK(cudaMemcpyAsync(sid, h.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, c->stream));
        rc = exact_lists(c, ns, c->sol.x0, c->sol.y0, c->sol.z0, c->sol.type, sid, g, true, c->p.dim == 2, c->sol.sb, ns, off, &ids);
        if (rc) return rc;
    }
    const long long total = off[ns];
    if (total > 0x7fffffffLL) return MPHX_ERR_UNSUPPORTED;
    ids.resize((size_t)std::max<long long>(total, 1));
    std::vector<int> off32((size_t)ns + 1);
    for (int s = 0; s <= ns; ++s) off32[s] = (int)off[s];
    // Store every row in the reference's list order: buckets jCX, jCY, jCZ ascending over the
    // stencil offsets (:1591-1620).  The reference configuration normally holds at most one solid particle per
    // bucket (lattice at cell spacing); if a bucket holds several, the reference orders them by its
    // (unstable) bitonic sort -- ties fall back to id order here and the solid path is then only
    // guaranteed to 1e-10, not bit for bit (reported through mphx_get_status).
    {
        std::vector<double> hx0(ns), hy0(ns), hz0(ns);
        CK(cudaMemcpy(hx0.data(), c->sol.x0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hy0.data(), c->sol.y0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hz0.data(), c->sol.z0, sizeof(double) * ns, cudaMemcpyDeviceToHost));
        auto coord = [](double x, double mn, double cw, int n) {
            int v = ((int)std::floor((x - mn) / cw)) % n;
            return (v % n + n) % n;
        };
        std::vector<int> cx(ns), cy(ns), cz(ns);
        std::vector<long long> cell(ns);
        for (int s = 0; s < ns; ++s) {
            cx[s] = coord(hx0[s], g.mn[0], g.cellw, g.nx);
            cy[s] = coord(hy0[s], g.mn[1], g.cellw, g.ny);
            cz[s] = g.dim == 3 ? coord(hz0[s], g.mn[2], g.cellw, g.nz) : 0;
            cell[s] = ((long long)cx[s] * g.ny + cy[s]) * g.nz + cz[s];
        }
        {
            std::vector<long long> sorted(cell);
            std::sort(sorted.begin(), sorted.end());
            c->solid_multi_occupancy = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
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
    // split sub-steps (slab mode): which ranks' rows read solid j -- rank(i) for every i whose own row lists j (pass 1
    // gathers u_j) or whose transposed row lists j (pass 2 gathers P_j); the rank advancing j itself needs no copy
    if (c->slab && c->split_substeps) {
        auto rank_of = [&](int s) { // inverse of s_lo = ns * r / nranks
            int r = (int)(((long long)s * c->nranks + c->nranks - 1) / ns);
            r = std::min(std::max(r, 0), c->nranks - 1);
            while (r > 0 && s < (int)((long long)ns * r / c->nranks)) --r;
            while (r < c->nranks - 1 && s >= (int)((long long)ns * (r + 1) / c->nranks)) ++r;
            return r;
        };
        std::vector<int> rk((size_t)ns);
        for (int s = 0; s < ns; ++s) rk[s] = rank_of(s);
        std::vector<unsigned short> pm((size_t)ns, 0);
        for (int i = 0; i < ns; ++i) {
            for (int q = off32[i]; q < off32[i + 1]; ++q) pm[ids[q]] |= (unsigned short)(1u << rk[i]);
            for (int q = roff[i]; q < roff[i + 1]; ++q) pm[rnbr[q]] |= (unsigned short)(1u << rk[i]);
        }
        for (int s = 0; s < ns; ++s) pm[s] &= (unsigned short)~(1u << rk[s]);
        if (c->alloc(&c->sol.pmask, (size_t)ns)) return MPHX_ERR_NOMEM;
        CK(cudaMemcpy(c->sol.pmask, pm.data(), sizeof(unsigned short) * (size_t)ns, cudaMemcpyHostToDevice));
    }
    int e = 0;
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
    // The static pair data as a dictionary (see Solid): the sub-steps stream 6 instead of 36 bytes per pair.
    c->sol.packed = 0;
    if (!std::getenv("MPHX_SOLID_RAW") && total > 0) {
        std::vector<int> len(ns), rlen(ns);
        CK(cudaMemcpy(len.data(), c->sol.len, sizeof(int) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rlen.data(), c->sol.rlen, sizeof(int) * ns, cudaMemcpyDeviceToHost));
        int maxlen = 0, rmaxlen = 0;
        for (int s = 0; s < ns; ++s) { maxlen = std::max(maxlen, len[s]); rmaxlen = std::max(rmaxlen, rlen[s]); }
        struct Tuple { unsigned long long v[4]; bool operator==(const Tuple &o) const { return !std::memcmp(v, o.v, sizeof(v)); } };
        struct Hash { size_t operator()(const Tuple &t) const { unsigned long long h = 1469598103934665603ull; for (auto x : t.v) { h ^= x; h *= 1099511628211ull; h ^= h >> 29; } return (size_t)h; } };
        std::unordered_map<Tuple, int, Hash> dict;
        std::vector<Rec> table;
        bool ok = true;
        auto pack = [&](const double *dx, const double *dy, const double *dz, const double *dw, const std::vector<int> &ln, int ml,
                        std::vector<unsigned short> &out) {
            const size_t cnt = (size_t)ml * ns;
            std::vector<double> hx(cnt), hy(cnt), hz(cnt), hw(cnt);
            if (cudaMemcpy(hx.data(), dx, sizeof(double) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(hy.data(), dy, sizeof(double) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(hz.data(), dz, sizeof(double) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(hw.data(), dw, sizeof(double) * cnt, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; return; }
            out.assign(cnt, 0);
            for (int s = 0; s < ns && ok; ++s)
                for (int kk = 0; kk < ln[s]; ++kk) {
                    const size_t q = (size_t)kk * ns + s;
                    Tuple t;
                    std::memcpy(&t.v[0], &hx[q], 8); std::memcpy(&t.v[1], &hy[q], 8); std::memcpy(&t.v[2], &hz[q], 8); std::memcpy(&t.v[3], &hw[q], 8);
                    auto it = dict.find(t);
                    int idx;
                    if (it == dict.end()) {
                        idx = (int)table.size();
                        if (idx > 65535) { ok = false; break; }
                        dict.emplace(t, idx);
                        Rec r; r.a = hx[q]; r.b = hy[q]; r.c = hz[q]; r.d = hw[q];
                        table.push_back(r);
                    } else idx = it->second;
                    out[q] = (unsigned short)idx;
                }
        };
        std::vector<unsigned short> tix, rtix;
        pack(c->sol.d0x, c->sol.d0y, c->sol.d0z, c->sol.w, len, maxlen, tix);
        if (ok) pack(c->sol.rd0x, c->sol.rd0y, c->sol.rd0z, c->sol.rw, rlen, rmaxlen, rtix);
        cudaGetLastError();
        if (ok) {
            int e2 = 0;
            e2 |= c->alloc(&c->sol.tix, tix.size()); e2 |= c->alloc(&c->sol.rtix, rtix.size()); e2 |= c->alloc(&c->sol.ttab, table.size());
            if (e2) return MPHX_ERR_NOMEM;
            CK(cudaMemcpy(c->sol.tix, tix.data(), sizeof(unsigned short) * tix.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->sol.rtix, rtix.data(), sizeof(unsigned short) * rtix.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->sol.ttab, table.data(), sizeof(Rec) * table.size(), cudaMemcpyHostToDevice));
            c->sol.packed = 1;
            c->solid_tuples = (int)table.size();
            // the same indices in CSR order for the team kernels (a team's lanes read consecutive entries of a row)
            std::vector<unsigned short> ctix((size_t)std::max<long long>(total, 1)), crtix((size_t)std::max<long long>(total, 1));
            for (int s = 0; s < ns; ++s) {
                for (int kk = 0; kk < len[s]; ++kk) ctix[(size_t)off32[s] + kk] = tix[(size_t)kk * ns + s];
                for (int kk = 0; kk < rlen[s]; ++kk) crtix[(size_t)roff[s] + kk] = rtix[(size_t)kk * ns + s];
            }
            if (c->alloc(&c->sol.ctix, ctix.size()) || c->alloc(&c->sol.crtix, crtix.size())) return MPHX_ERR_NOMEM;
            CK(cudaMemcpy(c->sol.ctix, ctix.data(), sizeof(unsigned short) * ctix.size(), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(c->sol.crtix, crtix.data(), sizeof(unsigned short) * crtix.size(), cudaMemcpyHostToDevice));
            c->sol_maxlen = maxlen; c->sol_rmaxlen = rmaxlen;
            if (c->solid_team) {
                c->team_ok = true;
            }
        }
    }
    return MPHX_OK;
}

// the half step of mphx_init: buckets, candidate list and the density sums on the initial positions (:565-568)
static int init_enqueue(Ctx *c)
{
    int rc;
    if ((rc = stage_build(c, false))) return rc;
    return run_pass1(c);
}

} // namespace mphx

using namespace mphx;

// =================================================================================================
extern "C" {

static int slab_check_errors(Ctx *c, const char *where);
static int slab_allocate(Ctx *c);

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
    if (p->clamp_module < 0 || p->clamp_module > MPHX_MODULE_ROLLING2) return MPHX_ERR_UNSUPPORTED;
    if (p->wall_module != MPHX_WALL_DEFAULT && p->wall_module != MPHX_WALL_ROLLING) return MPHX_ERR_UNSUPPORTED;
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
    if (const char *e = std::getenv("MPHX_LIST_REUSE")) c->list_reuse = std::atoi(e) != 0; // 0: rebuild buckets + list every step
    if (const char *e = std::getenv("MPHX_LIST_SKIN")) c->skin = std::max(0.0, std::atof(e)); // Verlet skin in particle spacings
    int rc = setup_constants(c);
    if (rc) { delete c; return rc; }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_last_error("cudaSetDevice/cudaStreamCreate failed");
        delete c;
        return MPHX_ERR_CUDA;
    }
    preload_kernels(p->dim);
    {
        Ctl init{};
        init.force = 1;
        if (c->alloc(&c->ctl, 1) || cudaMemcpy(c->ctl, &init, sizeof(Ctl), cudaMemcpyHostToDevice) != cudaSuccess) {
            mphx_destroy(reinterpret_cast<mphx_ctx *>(c));
            return MPHX_ERR_NOMEM;
        }
    }
    c->timing = std::getenv("MPHX_TIMING") != nullptr;
    if (const char *e = std::getenv("MPHX_SWEEP_BATCH")) c->sweep_batch = std::max(1, std::atoi(e));
    if (const char *e = std::getenv("MPHX_LIST_CAP")) c->list_cap = std::max(0, std::atoi(e)); // 0: no list, fused sweeps
    if (const char *e = std::getenv("MPHX_FILTER2")) c->filter2 = std::atoi(e) != 0;           // 0: one particle per thread
    if (const char *e = std::getenv("MPHX_BRICK")) c->brick = std::atoi(e) != 0;               // 1: pass 1 staged in shared memory (3D)
    cudaFuncSetAttribute(k_brick_pass1<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kBrickCap * (sizeof(Rec) + sizeof(double2))));
    for (int k = 0; k < 2; ++k) cudaEventCreate(&c->tev[k]);
    if (const char *e = std::getenv("MPHX_OVERLAP_SOLID")) c->overlap_solid = std::atoi(e) != 0;
    if (const char *e = std::getenv("MPHX_SOLID_TEAM")) { // 0: one thread per solid in the sub-steps, 2: a team per solid even on a full set
        c->solid_team = std::atoi(e) != 0;
        c->solid_team_always = std::atoi(e) == 2;
    }
    if (const char *e = std::getenv("MPHX_RING_INKERNEL_WAIT")) c->ring_inkernel_wait = std::atoi(e) != 0;
    if (const char *e = std::getenv("MPHX_SPLIT_SUBSTEPS")) c->split_substeps = std::atoi(e) != 0; // 0: every slab runs all solids' sub-steps
    {   // highest priority: the few blocks of a sub-step kernel must get SM slots as pass-2 blocks retire,
        // not after the whole pass-2 grid has been issued
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi) != cudaSuccess) c->side = nullptr;
    }
    cudaEventCreateWithFlags(&c->ev_solid_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_solid_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_u_ready, cudaEventDisableTiming);
    if (const char *e = std::getenv("MPHX_EARLY_PASS1")) c->early_pass1 = std::atoi(e) != 0;
    *out = reinterpret_cast<mphx_ctx *>(c);
    return MPHX_OK;
}

void mphx_destroy(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->side_marks) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (int r = 0; r < kMaxRanks; ++r)
        if (c->peer_ipc[r] && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    for (int k = 0; k < 2; ++k) if (c->tev[k]) cudaEventDestroy(c->tev[k]);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->ev_solid_ready) cudaEventDestroy(c->ev_solid_ready);
    if (c->ev_solid_done) cudaEventDestroy(c->ev_solid_done);
    if (c->ev_u_ready) cudaEventDestroy(c->ev_u_ready);
    for (void *q : c->allocs) cudaFree(q);
    if (c->stream && !c->external_stream) cudaStreamDestroy(c->stream);
    delete c;
}

// device allocations of the first upload (slots, lists, buckets, solid arrays) for `nloc` local particles of `n`
static int upload_allocate(Ctx *c, int n, int nloc, const int r[6])
{
    const bool first = !c->uploaded;
    if (first) {
        c->n_global = n;
        if (!c->slab) c->cap = n;
        c->nmax = c->cap;
        if (nloc > c->cap) { set_last_error("slab capacity too small for the initial particle set"); return MPHX_ERR_NOMEM; }
        const size_t cap = (size_t)c->cap;
        std::memcpy(c->ranges, r, sizeof(int) * 6);
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
        if (c->brick && c->pl.nbr && c->p.dim == 3 && !c->surface_tension && c->grid.range <= kBrickMaxRange) {
            c->nbricks = brick_grid(c->grid).nbricks;
            e |= c->alloc(&c->lnbr, ((size_t)c->pl.L + 1) * cap);
            e |= c->alloc(&c->brick_ok, (size_t)c->nbricks); e |= c->alloc(&c->in_brick, cap);
            c->brick_scan_blocks = (c->nbricks + kScanChunk - 1) / kScanChunk;
            e |= c->alloc(&c->brick_nown, (size_t)c->nbricks); e |= c->alloc(&c->brick_base, (size_t)c->nbricks + 1);
            e |= c->alloc(&c->brick_sums, (size_t)c->brick_scan_blocks + 1);
            if (!e) { cudaMemset(c->brick_ok, 0, (size_t)c->nbricks); cudaMemset(c->in_brick, 0, cap); }
        }
        c->cells_cap = (size_t)c->grid.ncells;
        e |= c->alloc(&c->cellCount, c->cells_cap + 2);
        e |= c->alloc(&c->cellStart, c->cells_cap + 3);
        c->scan_blocks = (int)(((long long)c->grid.ncells + 2 + kScanChunk - 1) / kScanChunk);
        e |= c->alloc(&c->blockSums, (size_t)c->scan_blocks + 1);
        e |= c->alloc(&c->slot, cap); e |= c->alloc(&c->tmpIdx, cap); e |= c->alloc(&c->where, cap);
        e |= c->alloc(&c->P, cap); e |= c->alloc(&c->volStrain, cap); e |= c->alloc(&c->divP, cap);
        e |= c->alloc(&c->fx, cap); e |= c->alloc(&c->fy, cap); e |= c->alloc(&c->fz, cap);
        e |= c->alloc(&c->ax, cap); e |= c->alloc(&c->ay, cap); e |= c->alloc(&c->az, cap);
        e |= c->alloc(&c->densA, cap); e |= c->alloc(&c->gcx, cap); e |= c->alloc(&c->gcy, cap);
        e |= c->alloc(&c->gcz, cap); e |= c->alloc(&c->PA, cap);
        e |= c->alloc(&c->d_inv_density, kTypeCount);
        e |= c->alloc(&c->ancx, cap); e |= c->alloc(&c->ancy, cap); e |= c->alloc(&c->ancz, cap);
        Solid &so = c->sol;
        so.ns = c->ns; so.sb = c->ns > 0 ? r[2] : 0;
        const size_t ns = (size_t)c->ns;
        c->s_lo = 0; c->s_hi = c->ns;
        if (c->slab) { // the mailbox first: the solid arrays other ranks store into (x, v, u, P) live inside it
            int src = slab_allocate(c);
            if (src) return src;
            so.x = c->mine.sxv[0]; so.y = c->mine.sxv[1]; so.z = c->mine.sxv[2];
            so.vx = c->mine.sxv[3]; so.vy = c->mine.sxv[4]; so.vz = c->mine.sxv[5];
            so.u = c->mine.su; so.PkA = c->mine.sPk;
            if (c->split_substeps) { // equal shares of the static solid order
                c->s_lo = (int)((long long)c->ns * c->rank / c->nranks);
                c->s_hi = (int)((long long)c->ns * (c->rank + 1) / c->nranks);
            }
        } else {
            // (developer timing hook: advance only the first ns / k solids, as one rank of k would -- wrong physics, right kernel time)
            if (const char *e2 = std::getenv("MPHX_DEBUG_SUBSTEP_SHARE")) c->s_hi = c->ns / std::max(1, std::atoi(e2));
            double **sx[] = {&so.x, &so.y, &so.z, &so.vx, &so.vy, &so.vz};
            for (double **q : sx) e |= c->alloc(q, ns);
            e |= c->alloc(&so.u, ns);
            e |= c->alloc(&so.PkA, 9 * ns);
        }
        double **sv[] = {&so.x0, &so.y0, &so.z0, &so.fx, &so.fy, &so.fz, &so.lam, &so.mu};
        for (double **q : sv) e |= c->alloc(q, ns);
        double **st[] = {&so.Linv, &so.Fm, &so.E, &so.S};
        for (double **q : st) e |= c->alloc(q, 9 * ns);
        e |= c->alloc(&so.type, ns); e |= c->alloc(&so.slot, ns);
        if (e) return MPHX_ERR_NOMEM;
        CK(cudaMemcpy(c->d_inv_density, c->phys.inv_density, sizeof(double) * kTypeCount, cudaMemcpyHostToDevice));
        CK(cudaMemsetAsync(c->cellCount, 0, sizeof(int) * ((size_t)c->grid.ncells + 2), c->stream)); // (all-zero outside a rebuild)
        double *zs[] = {c->P, c->volStrain, c->divP, c->fx, c->fy, c->fz, c->ax, c->ay, c->az, c->densA, c->gcx, c->gcy, c->gcz, c->PA,
                        c->ancx, c->ancy, c->ancz};
        for (double *q : zs) CK(cudaMemsetAsync(q, 0, sizeof(double) * cap, c->stream));
        if (ns > 0) {
            for (double **q : st) CK(cudaMemsetAsync(*q, 0, sizeof(double) * 9 * ns, c->stream));
            CK(cudaMemsetAsync(so.PkA, 0, sizeof(double) * 9 * ns, c->stream));
        }
    }
    return MPHX_OK;
}

// device staging arrays (original order, AoS) -> the context's slots and solid arrays; frees nothing
static int upload_finish(Ctx *c, int nloc, const int *d_ids, const int *d_t, const double *d_x, const double *d_x0, const double *d_v)
{
    if (nloc > 0) LAUNCH(c, k_upload_split, nblk(nloc), kBlock, nloc, d_ids, d_t, d_x, d_v, c->S, c->sol.slot, c->sol.sb);
    if (c->ns > 0) LAUNCH(c, k_solid_upload, nblk(c->ns), kBlock, c->sol, d_t, d_x, d_x0, d_v);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    CK(cudaMemcpy(&c->ctl->n, &nloc, sizeof(int), cudaMemcpyHostToDevice));
    c->uploaded = true;
    c->bx = c->S.x; c->by = c->S.y; c->bz = c->S.z;
    { int rrc = request_rebuild(c); if (rrc) return rrc; } // (state replaced: the next step rebuilds buckets and list)
    CK(cudaStreamSynchronize(c->stream));
    return MPHX_OK;
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
    { int arc = upload_allocate(c, n, nloc, r); if (arc) return arc; }
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
    const int frc = upload_finish(c, nloc, d_ids, d_t, d_x, d_x0, d_v);
    cudaFree(d_t); cudaFree(d_x); cudaFree(d_v); cudaFree(d_x0);
    if (d_ids) cudaFree(d_ids);
    return frc;
}

// ---- device-side generator --------------------------------------------------------------
// value after a printf("%e") / strtod round trip: what the solver reads from the generator's .grid text
static double through_e(double v)
{
    char b[64];
    std::snprintf(b, sizeof(b), "%e", v);
    return std::strtod(b, nullptr);
}
// one axis of generator/generator.cpp:654-680: start at lower + spacing/2, accumulate `p += spacing` while p < upper - 0.49 spacing
static void generator_axis(double lo, double hi, double space, std::vector<double> &out)
{
    const double width = hi - lo;
    const int count = (int)std::round(width / space);
    const double spacing = width / count;
    for (double p = lo + 0.5 * spacing; p < hi - 0.49 * spacing; p += spacing) out.push_back(through_e(p));
}
static int generator_plan(const mphx_cuboid *cubs, int ncub, GenPlan &plan, std::vector<double> &axes, long long &total, int r[6])
{
    if (!cubs || ncub < 1 || ncub > kMaxCuboids) return MPHX_ERR_INVALID;
    plan.count = ncub;
    total = 0;
    for (int k = 0; k < 6; ++k) r[k] = -1;
    int last_cls = -1;
    bool seen[3] = {false, false, false};
    for (int q = 0; q < ncub; ++q) {
        const mphx_cuboid &cb = cubs[q];
        if (cb.type < 0 || cb.type >= kTypeCount || !(cb.spacing > 0.0)) return MPHX_ERR_INVALID;
        GenCuboid &g = plan.c[q];
        g.first = total; g.type = cb.type;
        int cnt[3];
        for (int d = 0; d < 3; ++d) {
            const size_t before = axes.size();
            generator_axis(cb.lower[d], cb.upper[d], cb.spacing, axes);
            cnt[d] = (int)(axes.size() - before);
            (d == 0 ? g.ax : d == 1 ? g.ay : g.az) = (int)before;
            g.v[d] = through_e(cb.velocity[d]);
        }
        g.nx = cnt[0]; g.ny = cnt[1]; g.nz = cnt[2];
        const long long np = (long long)cnt[0] * cnt[1] * cnt[2];
        if (np <= 0) continue;
        const int cls = cb.type < 2 ? 0 : cb.type < 4 ? 1 : 2;
        // each class must be one contiguous block of the file order (src/main.cpp:909-929)
        if (cls != last_cls && seen[cls]) { set_last_error("particle classes are not contiguous in cuboid order"); return MPHX_ERR_UNSUPPORTED; }
        seen[cls] = true; last_cls = cls;
        if (r[2 * cls] < 0) r[2 * cls] = (int)total;
        total += np;
        r[2 * cls + 1] = (int)total;
    }
    if (total <= 0 || total > 0x7fffffffLL) return MPHX_ERR_INVALID;
    return MPHX_OK;
}

long long mphx_generate_count(const mphx_cuboid *cuboids, int ncuboids)
{
    GenPlan plan;
    std::vector<double> axes;
    long long total = 0;
    int r[6];
    return generator_plan(cuboids, ncuboids, plan, axes, total, r) == MPHX_OK ? total : -1;
}

// owned-by-column histogram of the fluid / wall particles the cuboids hold, from the axis tables alone (no particle arrays):
// what cuts the slabs of a generated case (mphx_partition_columns)
int mphx_generate_column_histogram(const mphx_cuboid *cuboids, int ncuboids, double domain_min0, double cell_width, int ncols, long long *hist)
{
    if (!hist || ncols < 1 || !(cell_width > 0.0)) return MPHX_ERR_INVALID;
    GenPlan plan;
    std::vector<double> axes;
    long long total = 0;
    int r[6];
    int rc = generator_plan(cuboids, ncuboids, plan, axes, total, r);
    if (rc) return rc;
    for (int i = 0; i < ncols; ++i) hist[i] = 0;
    for (int q = 0; q < plan.count; ++q) {
        const GenCuboid &g = plan.c[q];
        if (g.type >= 2 && g.type < 4) continue;
        for (int ix = 0; ix < g.nx; ++ix) {
            int cx = ((int)std::floor((axes[(size_t)g.ax + ix] - domain_min0) / cell_width)) % ncols; // :1671
            cx = (cx % ncols + ncols) % ncols;
            hist[cx] += (long long)g.ny * g.nz;
        }
    }
    return MPHX_OK;
}

int mphx_upload_generated(mphx_ctx *ctx, const mphx_cuboid *cuboids, int ncuboids)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || c->uploaded) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    GenPlan plan;
    std::vector<double> axes;
    long long total = 0;
    int r[6];
    int rc = generator_plan(cuboids, ncuboids, plan, axes, total, r);
    if (rc) return rc;
    const int n = (int)total;
    Scratch tmp;
    int *d_t;
    double *d_x, *d_x0, *d_v, *d_axes;
    if (tmp.get(&d_t, (size_t)n) || tmp.get(&d_x, 3 * (size_t)n) || tmp.get(&d_x0, 3 * (size_t)n) || tmp.get(&d_v, 3 * (size_t)n) ||
        tmp.get(&d_axes, axes.size()))
        return MPHX_ERR_NOMEM;
    CK(cudaMemcpyAsync(d_axes, axes.data(), sizeof(double) * axes.size(), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_generate, nblk(n), kBlock, (long long)n, plan, d_axes, d_t, d_x, d_x0, d_v);
    CK(cudaGetLastError());
    if (!c->slab) {
        if ((rc = upload_allocate(c, n, n, r))) return rc;
        return upload_finish(c, n, nullptr, d_t, d_x, d_x0, d_v);
    }
    // a slab keeps all solids and the fluid / wall particles of its own columns: mask, scan, ids -- on the device
    int *keep, *scan, *sums, *d_ids;
    const int sb = (int)(((long long)n + 1 + kScanChunk - 1) / kScanChunk);
    if (tmp.get(&keep, (size_t)n + 1) || tmp.get(&scan, (size_t)n + 2) || tmp.get(&sums, (size_t)sb + 1)) return MPHX_ERR_NOMEM;
    CK(cudaMemsetAsync(keep + n, 0, sizeof(int), c->stream));
    LAUNCH(c, k_generated_keep, nblk(n), kBlock, (long long)n, (const int *)d_t, (const double *)d_x, c->grid, keep);
    LAUNCH(c, k_scan_reduce, sb, kScanThreads, (const Ctl *)nullptr, (const int *)keep, n + 1, sums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, (const Ctl *)nullptr, sums, sb);
    LAUNCH(c, k_scan_apply, sb, kScanThreads, (const Ctl *)nullptr, keep, n + 1, (const int *)sums, scan, 0);
    int nloc = 0;
    CK(cudaMemcpyAsync(&nloc, scan + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if ((rc = upload_allocate(c, n, nloc, r))) return rc;
    if (tmp.get(&d_ids, (size_t)std::max(nloc, 1))) return MPHX_ERR_NOMEM;
    LAUNCH(c, k_generated_ids, nblk(n), kBlock, (long long)n, (const int *)keep, (const int *)scan, d_ids);
    CK(cudaGetLastError());
    return upload_finish(c, nloc, d_ids, d_t, d_x, d_x0, d_v);
}

// replace Position and Velocity of every particle (original order) on an initialised context:
// the per-step host->device path of a caller that owns the state on the host
int mphx_upload_state(mphx_ctx *ctx, const double *position, const double *velocity)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !position || !velocity) return MPHX_ERR_INVALID;
    if (c->slab) { set_last_error("mphx_upload_state: use mphx_upload_owned on a slab context"); return MPHX_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    const size_t N = (size_t)c->n_global;
    if (!c->stage3a && c->alloc(&c->stage3a, 3 * N)) return MPHX_ERR_NOMEM;
    if (!c->stage3b && c->alloc(&c->stage3b, 3 * N)) return MPHX_ERR_NOMEM;
    CK(cudaMemcpyAsync(c->stage3a, position, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->stage3b, velocity, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_upload_state, nblk(c->nmax), kBlock, c->nmax, c->S, c->sol, c->stage3a, c->stage3b);
    c->bx = c->S.x; c->by = c->S.y; c->bz = c->S.z;
    CK(cudaGetLastError());
    return request_rebuild(c); // arbitrary new positions: the next step rebuilds buckets and list
}

int mphx_init(mphx_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->uploaded) return MPHX_ERR_INVALID;
    if (c->inited) return MPHX_OK;
    if (c->slab && !c->connected) { set_last_error("mphx_init: connect the slab to its peers first (mphx_slab_connect)"); return MPHX_ERR_INVALID; }
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = init_solid(c))) return rc; // calculateInitialNeighbor, Lame, Normalizer
    // first calculateNeighbor + density sums on the initial positions (:565-568): gives
    // NeighborCount / PressureP for the `output.vtk` written before the loop (:572).
    // (slab mode: every rank must be in this call at the same time -- the halos are exchanged)
    if ((rc = init_enqueue(c))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->inited = true;
    return slab_check_errors(c, "mphx_init");
}

int mphx_get_constants(const mphx_ctx *ctx, mphx_constants *k)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !k) return MPHX_ERR_INVALID;
    *k = c->c;
    return MPHX_OK;
}

int mphx_get_wall_centers(const mphx_ctx *ctx, double centers[MPHX_TYPE_COUNT][3])
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !centers) return MPHX_ERR_INVALID;
    std::memcpy(centers, c->wall_center, sizeof(double) * 3 * MPHX_TYPE_COUNT);
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
    return slab_check_errors(c, "mphx_sync");
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
    int n = 0;
    { int nrc = read_n(c, &n); if (nrc) return nrc; }
    const int ns = c->ns;
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
    // owner_only: the solids' values live on the slab that evaluates them (Force in slab mode), not replicated
    auto vec3 = [&](double *host, const double *a, const double *b, const double *cc, const double *sa, const double *sb_, const double *sc,
                    bool owner_only = false) -> int {
        if (!host) return MPHX_OK;
        owner_only = owner_only && c->slab;
        if (c->slab) CK(cudaMemsetAsync(d3, 0, sizeof(double) * 3 * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_vec3, nblk(n), kBlock, n, S.id, c->mask, a, b, cc, d3);
        if (ns > 0 && sa && (report_solids || owner_only))
            LAUNCH(c, k_solid_vec3_to_orig, nblk(ns), kBlock, so, sa, sb_, sc, d3, owner_only ? (const int *)S.type : (const int *)nullptr);
        CK(cudaMemcpyAsync(host, d3, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, c->stream));
        return MPHX_OK; // (no synchronisation per field: the staging buffer is reused in stream order)
    };
    auto scal = [&](double *host, const double *a) -> int {
        if (!host) return MPHX_OK;
        if (c->slab) CK(cudaMemsetAsync(d1, 0, sizeof(double) * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_scalar, nblk(n), kBlock, n, S.id, c->mask2, a, d1);
        CK(cudaMemcpyAsync(host, d1, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        return MPHX_OK;
    };
    auto ints = [&](int *host, const int *a, const int *msk) -> int {
        if (!host) return MPHX_OK;
        if (c->slab) CK(cudaMemsetAsync(di, 0, sizeof(int) * N, c->stream));
        if (n > 0) LAUNCH(c, k_gather_int, nblk(n), kBlock, n, S.id, msk, a, di);
        CK(cudaMemcpyAsync(host, di, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        return MPHX_OK;
    };
    auto tens = [&](double *host, const double *M) -> int {
        if (!host) return MPHX_OK;
        if (!d9) {
            if (c->alloc(&c->stage9, 9 * N)) return MPHX_ERR_NOMEM;
            d9 = c->stage9;
        }
        CK(cudaMemsetAsync(d9, 0, sizeof(double) * 9 * N, c->stream));
        // split sub-steps: F, E, S of a solid live on the rank that advances it (the sum over the slabs is the case)
        const bool ranged = c->slab && c->split_substeps;
        const int lo = ranged ? c->s_lo : 0, hi = ranged ? c->s_hi : ns;
        if (ns > 0 && (report_solids || ranged)) LAUNCH(c, k_solid_tensor_to_orig, nblk(hi - lo), kBlock, so, M, d9, lo, hi);
        CK(cudaMemcpyAsync(host, d9, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost, c->stream));
        return MPHX_OK;
    };
    auto solid_scal = [&](double *host, const double *a) -> int {
        if (!host) return MPHX_OK;
        CK(cudaMemsetAsync(d1, 0, sizeof(double) * N, c->stream));
        if (ns > 0 && report_solids) LAUNCH(c, k_solid_scalar_to_orig, nblk(ns), kBlock, so, a, d1);
        CK(cudaMemcpyAsync(host, d1, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        return MPHX_OK;
    };
    do {
        if (v->property) {
            if (n > 0) LAUNCH(c, k_real_types, nblk(n), kBlock, n, S, c->tmpi);
            if ((rc = ints(v->property, c->tmpi, c->mask))) break;
        }
        if ((rc = vec3(v->position, S.x, S.y, S.z, so.x, so.y, so.z))) break;
        if ((rc = vec3(v->velocity, S.vx, S.vy, S.vz, so.vx, so.vy, so.vz))) break;
        if ((rc = vec3(v->force, c->fx, c->fy, c->fz, so.fx, so.fy, so.fz, true))) break;
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
            std::vector<long long> off;
            if ((rc = exact_lists_current(c, off, nullptr))) break;
            for (size_t i = 0; i < N; ++i) v->neighbor_count[i] = (int)(off[i + 1] - off[i]);
        }
        if (v->initial_structure_neighbor_count) {
            CK(cudaMemsetAsync(di, 0, sizeof(int) * N, c->stream));
            if (ns > 0 && so.off && report_solids) LAUNCH(c, k_solid_rowlen_to_orig, nblk(ns), kBlock, so, di);
            CK(cudaMemcpyAsync(v->initial_structure_neighbor_count, di, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
        }
        if ((rc = tens(v->normalizer, so.Linv))) break;
        if ((rc = tens(v->deform_gradient, so.Fm))) break;
        if ((rc = tens(v->strain, so.E))) break;
        if ((rc = tens(v->stress, so.S))) break;
        if ((rc = solid_scal(v->lambda_lames, so.lam))) break;
        if ((rc = solid_scal(v->mu_lames, so.mu))) break;
        if (v->virial_stress || v->virial_pressure) { // calculateVirialStressAtParticle :3077-3318 (N2)
            if (!c->inited) { rc = MPHX_ERR_INVALID; break; }
            if (!d9) {
                if (c->alloc(&c->stage9, 9 * N)) return MPHX_ERR_NOMEM;
                d9 = c->stage9;
            }
            cudaEvent_t t0 = timer_event(c), t1 = timer_event(c);
            cudaEventRecord(t0, c->stream);
            if ((rc = compute_virial(c, d9, d1))) break;
            cudaEventRecord(t1, c->stream);
            cudaEventSynchronize(t1);
            float vms = 0.f;
            cudaEventElapsedTime(&vms, t0, t1);
            c->virial_ms += vms;
            c->ev_pool.push_back(t0); c->ev_pool.push_back(t1);
            if (v->virial_stress) CK(cudaMemcpyAsync(v->virial_stress, d9, sizeof(double) * 9 * N, cudaMemcpyDeviceToHost, c->stream));
            if (v->virial_pressure) CK(cudaMemcpyAsync(v->virial_pressure, d1, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
        }
    } while (0);
    CK(cudaStreamSynchronize(c->stream)); // the one synchronisation of the call
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
    int n = 0;
    if ((rc = read_n(c, &n))) return rc;
    LAUNCH(c, k_owned_mask, nblk(n), kBlock, n, c->S, c->grid, 1, c->mask);
    const int nb = (int)(((long long)n + kScanChunk - 1) / kScanChunk);
    LAUNCH(c, k_scan_reduce, nb, kScanThreads, (const Ctl *)nullptr, c->mask, n, c->own_sums);
    LAUNCH(c, k_scan_top, 1, kScanThreads, (const Ctl *)nullptr, c->own_sums, nb);
    LAUNCH(c, k_scan_apply, nb, kScanThreads, (const Ctl *)nullptr, c->mask, n, c->own_sums, c->own_scan, 0);
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
    int *d_err = c->own_scan; // (scratch: free between a download and the next one)
    CK(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
    LAUNCH(c, k_scatter_owned, nblk(count), kBlock, count, c->S, c->sol, c->own_slot, c->own_ids, c->own_x, c->own_v, d_err);
    int err = 0;
    CK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (err) { set_last_error("mphx_upload_owned: ids do not match the rows of the last download"); return MPHX_ERR_INVALID; }
    c->bx = c->S.x; c->by = c->S.y; c->bz = c->S.z;
    CK(cudaGetLastError());
    return request_rebuild(c); // arbitrary new positions: the next step rebuilds buckets and list
}

int mphx_debug_neighbors(mphx_ctx *ctx, long long *offsets, int *ids, long long cap)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !offsets) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    std::vector<long long> off;
    std::vector<int> rows;
    int rc = exact_lists_current(c, off, ids ? &rows : nullptr);
    if (rc) return rc;
    std::memcpy(offsets, off.data(), sizeof(long long) * ((size_t)c->n_global + 1));
    if (ids) {
        if (cap < (long long)rows.size()) return MPHX_ERR_OVERFLOW;
        std::memcpy(ids, rows.data(), sizeof(int) * rows.size());
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

/* Device-side timeline of the following steps: capacity > 0 gives the context a buffer of that many (code, ns) marks and
   switches the markers on (a few one-thread kernels per step), 0 switches them off.  Codes: see trace_mark in kernels.cuh. */
int mphx_trace_enable(mphx_ctx *ctx, int capacity)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || capacity < 0) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    CK(cudaStreamSynchronize(c->stream));
    unsigned long long *buf = nullptr;
    if (capacity > 0) { // (the buffer is kept and reused by later calls)
        if (capacity > c->trace_cap_alloc) {
            if (c->alloc(&c->trace_buf, 2 * (size_t)capacity)) return MPHX_ERR_NOMEM;
            c->trace_cap_alloc = capacity;
        }
        buf = c->trace_buf;
    }
    const int zero = 0;
    CK(cudaMemcpy(&c->ctl->trace, &buf, sizeof(buf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&c->ctl->trace_cap, &capacity, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&c->ctl->trace_n, &zero, sizeof(int), cudaMemcpyHostToDevice));
    c->tracing = capacity > 0;
    return MPHX_OK;
}
/* the marks recorded so far: out[2 i] = code, out[2 i + 1] = %globaltimer in ns; *count = marks written (<= max_marks) */
int mphx_trace_read(mphx_ctx *ctx, unsigned long long *out, int max_marks, int *count)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !out || !count || max_marks < 0) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    CK(cudaStreamSynchronize(c->stream));
    Ctl h;
    CK(cudaMemcpy(&h, c->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    const int n = std::min(std::min(h.trace_n, h.trace_cap), max_marks);
    if (n > 0) CK(cudaMemcpy(out, h.trace, sizeof(unsigned long long) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
    *count = n;
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

/* dense FP64 FMA throughput of `device` in TFLOP/s, measured with CUDA events (best of 5 launches of a pure DFMA kernel) */
int mphx_measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0, 0));
        k_fp64_peak<<<blocks, threads>>>(d, iters);
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, tf);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    CK(cudaGetLastError());
    *tflops = best;
    return MPHX_OK;
}

/* candidates held by the current lists and pairs within the largest kernel radius, summed over the particles of this
 * context (out[0], out[1]); synchronises.  The algorithmic FP64 work of one sweep is ~15 flop per candidate + ~45 per
 * in-radius pair. */
int mphx_count_pairs(mphx_ctx *ctx, unsigned long long out[2])
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !c->inited || !out || !c->pl.nbr) return MPHX_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    { int jrc = join_solids(c); if (jrc) return jrc; }
    Scratch tmp;
    unsigned long long *d;
    if (tmp.get(&d, 2)) return MPHX_ERR_NOMEM;
    CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), c->stream));
    const double r = sweep_radius(c);
    LAUNCH(c, k_count_pairs, nblk(c->nmax), kBlock, c->ctl, c->S, c->grid, c->pl, r * r, d);
    CK(cudaMemcpyAsync(out, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return MPHX_OK;
}

/* accumulated device milliseconds of the virial stress diagnostic (the reference's "virial calculation" timer, :674) */
double mphx_get_virial_ms(const mphx_ctx *ctx) { return ctx ? reinterpret_cast<const Ctx *>(ctx)->virial_ms : 0.0; }

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
