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

// particle type as stored on the device: the reference's Property (0..5) in the low three bits; bit 3
// marks a ghost copy received from a neighbouring slab (never integrated, dropped every step)
constexpr int kGhost = 8;
// slab mode: this context evaluates the (replicated) solid -- it owned the solid's column when the current
// candidate list was built (ownership is frozen between list builds, like the lists themselves)
constexpr int kSolidOwned = 16;
__host__ __device__ inline int real_type(int t) { return t & 7; }
__host__ __device__ inline bool is_structure_type(int t) { return (t & 7) >= 2 && (t & 7) < 4; }
__host__ __device__ inline bool is_fluid_type(int t) { return (t & 7) >= 0 && (t & 7) < 2; }
__host__ __device__ inline bool is_wall_type(int t) { return (t & 7) >= 4 && (t & 7) < 6; }

struct GridDesc {
    int dim;        // 2 or 3
    int nx, ny, nz; // bucket counts (nz = 1 in 2D)
    int ncells;
    int nsten; // stencil columns
    int range; // ceil((MaxRadius+MARGIN)/CellWidth)  src/main.cpp:1744
    int slab;  // 1: this context holds one x-slab of a multi-GPU run (see slab.inc)
    // slab mode: nx is the LOCAL column count (owned columns [range, nx-range) + one halo of `range`
    // columns each side), xoff the global column of local column 0 (may be negative / wrap), nxg the
    // global column count, mn[0] the local origin mn0g + xoff*cellw.  Otherwise nxg=nx, xoff=0.
    int nxg, xoff;
    double mn0g;
    double mn[3], W[3], cellw;
    // stencil columns.  3D: (dx,dy) with half-length sh along z.  2D: dx with half-length sh along y.
    signed char sdx[kMaxStencil], sdy[kMaxStencil], sh[kMaxStencil];
};

// sticky error flags of a context (Ctl::err)
enum { kErrLost = 1,      // a particle crossed more than one halo width in a step
       kErrMsgFull = 2,   // exchange message buffer too small
       kErrArrival = 4,   // a received particle does not belong where it was sent
       kErrTimeout = 8,   // a peer's flag did not arrive (exchange wait timed out); bits 8.. say which: 256 votes, 512 migrants,
                          // 1024 halo, 2048 PressureP, 4096 solids' PressureP, 8192 solids' velocity
       kErrCapacity = 16, // particle slots exhausted
       kErrNaN = 32 };    // a non-finite position came out of the integration

// Device-resident control block of one context.  Everything the step decides -- how many slots are held,
// whether this step rebuilds the buckets and the candidate list or reuses them, the exchange counts -- lives
// here, so a step is a fixed sequence of kernel launches with no device->host read-back.
struct Ctl {
    int n;          // particle slots currently held (slab mode: owned + ghosts + all solids)
    int rebuild;    // this step: 1 = rebuild buckets + candidate list, 0 = reuse them
    int force;      // host request: rebuild at the next step (first step, uploads)
    int need;       // local reasons to rebuild this step (slab mode: OR-ed over the ranks)
    int age;        // steps the current list has served
    int last_age;   // steps the previous list served
    int probe;      // builds since the skin was last tried
    int skin_on;    // the current list carries the Verlet skin
    unsigned maxdisp2; // float bits: max |x - anchor|^2 over the particles this context moves, this step
    float filt2;    // squared filter radius (bucket units) of the list being built
    int err;        // kErr* bits
    int builds, reuses; // statistics
    int n_own_sol;  // slab mode: solids this context evaluates (list own_sol)
    int mig_cnt[2];   // emigrants packed for the left / right neighbour (this step)
    int halo_cnt[2];  // halo particles packed for left / right (current list)
    int ghost_base[2], ghost_cnt[2]; // pre-permute slots of the ghosts received from left / right (current list)
    unsigned push_done[8]; // block counters of the push kernels
    unsigned sub_done[2];  // block counters of the split solid sub-steps (pass 1 / pass 2)
    // device-side timeline (mphx_trace_*): (code, %globaltimer) pairs appended by the wait / push / marker kernels
    unsigned long long *trace;
    int trace_cap, trace_n;
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

struct __align__(8) PfPair { float2 x, y, z; }; // 24 bytes: x0 x1 y0 y1 z0 z1
struct __align__(32) Rec { double a, b, c, d; };

// cell-sorted particle arrays
struct Particles {
    double *x, *y, *z, *vx, *vy, *vz;
    int *type, *id, *key;
    // (position - DomainMin)/CellWidth in fp32, PAIR-interleaved (particles 2k, 2k+1 share one
    // 24-byte record x0 x1 y0 y1 z0 z1): input of the packed-fp32 filter.  Padded by two pairs.
    // (The filter is bound by the bytes the L1 returns to registers, so the record carries nothing else.)
    PfPair *pf;
    // 32-byte gather records of the sorted particles (what a NEIGHBOUR's thread reads per pair, one
    // 256-bit load each): ra[q] = (x, y, z, vx);  rb[q] = (vy, vz, PressureP, type bits).
    // Written by the permute (PressureP by pass 1); one buffer shared by both ping-pong sets.
    Rec *ra, *rb;
};

// total-Lagrangian solid, static order (solid-local index s = original id - sb)
struct Solid {
    int ns, sb;
    double *x, *y, *z, *vx, *vy, *vz, *x0, *y0, *z0, *fx, *fy, *fz;
    double *Linv, *Fm, *E, *S; // 9 planes of ns doubles each: M[k*ns+s], k = 3*row+col
    double *PkA;               // first Piola-Kirchhoff stress, 9 doubles per solid (AoS: a neighbour's thread gathers all nine)
    double *lam, *mu;
    int *type;
    int *off, *nbr;   // InitialStructureNeighbor as CSR (solid-local ids, rows in the reference's list order)
    int *roff, *rnbr; // transpose: rows = particles that list s (ascending)
    // The sub-step kernels read the lists in ELL layout (entry kk of row s at [kk*ns + s]: coalesced
    // across the threads of a warp) together with static per-pair data of the reference
    // configuration computed once: x0_ij and weight(x0_ij).
    int *len, *rlen, *rsplit;        // row lengths; rsplit = number of transposed entries of rows j < s
    int *slot;                       // sorted slot of solid s in the current bucket order (written by the permute)
    int *enbr, *ernbr;               // ELL neighbour ids (own rows / transposed rows)
    double *d0x, *d0y, *d0z, *w;     // own rows
    double *rd0x, *rd0y, *rd0z, *rw; // transposed: x0_js as row j computes it
    Rec *u;                          // displacement u = minimg(x - x0) of the current positions (x, y, z, -: one 256-bit gather)
    // Static pair data as a DICTIONARY: a lattice solid has a few thousand distinct (x0_ij, w) tuples among its
    // ~80 pairs per particle, so each pair stores a 16-bit index into `ttab` instead of four doubles (the sub-steps
    // are bound by streaming this data: 36 -> 6 bytes per pair).  packed = 0 (more than 65535 tuples): the raw arrays.
    int packed;
    unsigned short *tix, *rtix;      // ELL, own rows / transposed rows
    unsigned short *ctix, *crtix;    // the same indices in CSR order (off / roff): what the team kernels read
    Rec *ttab;                       // (x0_ij.x, .y, .z, weight(x0_ij))
    unsigned short *pmask;           // slab mode, split sub-steps: bit r = rank r's rows reference this solid (r != the rank advancing it)
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
// bucket key.  Buckets [0,ncells) are real; ncells = "parked" (a replicated solid outside this slab and
// its halo: kept, never traversed); ncells+1 = "dead" (ghosts of the previous step, particles that
// migrated away: dropped by the next permute).  *col receives the local x column.
__device__ __forceinline__ int cell_key(const GridDesc &g, double x, double y, double z, int *col = nullptr)
{
    int cx = cell_coord_exact(x, g.mn0g, g.cellw, g.nxg);
    if (g.slab) {
        cx -= g.xoff;
        if (cx < 0) cx += g.nxg;
        else if (cx >= g.nxg) cx -= g.nxg;
    }
    if (col) *col = cx;
    if (cx >= g.nx) return g.ncells;
    const int cy = cell_coord_exact(y, g.mn[1], g.cellw, g.ny);
    if (g.dim == 2) return cx * g.ny + cy;
    const int cz = cell_coord_exact(z, g.mn[2], g.cellw, g.nz);
    return (cx * g.ny + cy) * g.nz + cz;
}
__device__ __forceinline__ int key_column(const GridDesc &g, int key) { return g.dim == 2 ? key / g.ny : key / (g.ny * g.nz); }
__device__ __forceinline__ bool column_owned(const GridDesc &g, int cx) { return !g.slab || (cx >= g.range && cx < g.nx - g.range); }
// the reference's CellId (global) of a local key
__device__ __forceinline__ int global_key(const GridDesc &g, int key)
{
    if (!g.slab) return key;
    const int per = g.dim == 2 ? g.ny : g.ny * g.nz;
    int cx = key / per + g.xoff;
    if (cx < 0) cx += g.nxg;
    else if (cx >= g.nxg) cx -= g.nxg;
    return cx * per + key % per;
}
// plain (FMA-allowed) minimum image for the 1e-10 paths
__device__ __forceinline__ double minimg(double d, double W)
{
    const double h = 0.5 * W;
    const double t = d + h;
    return (t - W * floor(t / W)) - h;
}

// ------------------------------------------------------------------------------------------------
// Exchange between slabs (multi-GPU).  Every context owns one MAILBOX allocation that its ring
// neighbours (particles, PressureP) and all ranks (replicated solids, rebuild votes) write into with
// plain stores over NVLink -- peer memory mapped either directly (one process, several devices) or through
// CUDA IPC (one process per device).  A message is complete when its flag carries the step's epoch;
// receivers spin on the flag inside a one-warp kernel, so no count or flag ever travels through a host.
constexpr int kMaxRanks = 16;
constexpr int kMsgDoubles = 7; // x y z vx vy vz (type << 32 | id)
struct Mailbox {               // pointers into ONE context's mailbox (the layout is the same on every rank)
    unsigned long long *vote;  // [nranks] (epoch << 1 | need) : rebuild votes
    unsigned long long *fmig, *fhalo, *fp; // [2] each: from the left / right neighbour
    unsigned long long *fsolP, *fsolV;     // [nranks]
    int *cnt_mig, *cnt_halo;   // [2] each
    double *mig[2], *halo[2];  // [7 * msg_cap]
    double *p[2];              // [5][msg_cap]  PressureP (+ PressureA, GravityCenter) of the halo copies
    double *solP, *solV;       // [ns], [3 * ns]  replicated solids: PressureP and the coupled velocity
    // split solid sub-steps: the solid arrays other ranks store into live in the mailbox (Solid::x.., u, PkA point here)
    unsigned long long *fsub;  // [nranks] phase counter of every rank's sub-step kernels
    double *sxv[6];            // [ns] each: x y z vx vy vz
    Rec *su;                   // [ns] displacement records
    double *sPk;               // [9 * ns] first Piola-Kirchhoff stress
};
// the same arrays of every rank, as the sub-step kernels address them (the layout is identical on every rank)
struct SolidRing {
    int nranks, rank;          // nranks = 0: single context, nothing is pushed
    int last;                  // pass 2: last sub-step of the step -> positions and velocities go to every rank
    unsigned long long seq;    // value this launch posts into every rank's fsub[rank] when its last block is done
    unsigned long long wait_seq; // != 0: every block first waits until all ranks' fsub carry at least this value (the previous phase)
    char *base[kMaxRanks];
    size_t off_flag, off_xv[6], off_u, off_pk;
};
struct Peers {                 // the mailboxes a context writes into
    int nranks, rank;
    Mailbox left, right;       // ring neighbours
    double *solP[kMaxRanks], *solV[kMaxRanks];
    unsigned long long *fsolP[kMaxRanks], *fsolV[kMaxRanks], *vote[kMaxRanks];
};

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull; // a missing peer becomes an error flag, never a hang
// timeline marks (off unless mphx_trace_enable gave the context a buffer): 1..99 step phases (k_mark), 100 + tag / 200 + tag
// a wait begins / ends, 300 + which the last block of a solid sub-step kernel posts its phase, 400 + which a push completes
__device__ __forceinline__ void trace_mark(Ctl *ctl, int code)
{
    if (ctl->trace_cap > 0) {
        const int i = atomicAdd(&ctl->trace_n, 1);
        if (i < ctl->trace_cap) { ctl->trace[2 * i] = (unsigned long long)code; ctl->trace[2 * i + 1] = global_ns(); }
    }
}
__global__ void k_mark(Ctl *ctl, int code) { trace_mark(ctl, code); }

// wait until flags[0..nflags) carry this step's epoch (SHIFT = 1: the flag also carries a vote bit, the
// OR of the votes goes to ctl->need)
// (the epoch is a kernel argument, not device state: the waits of the solid sub-steps run on a second stream
// while the context's stream may already be enqueuing the next step)
enum { kWaitVote = 0, kWaitMig, kWaitHalo, kWaitP, kWaitSolP, kWaitSolV, kWaitSub };
constexpr int kSubPhases = 4096; // phase counters of the split solid sub-steps: epoch * kSubPhases + phase
struct DecideArgs {
    int reuse_enabled; // host switch (MPHX_LIST_REUSE, moving walls, slab support)
    float half_skin2;  // (skin / 2)^2 in metres^2
    float filt2_plain, filt2_skin;
};
__device__ __forceinline__ void decide_body(Ctl *ctl, const DecideArgs &a, int *__restrict__ list_flags);
// (flags2 / nflags2: a second set of flags waited for by the following lanes, e.g. the halo's PressureP and the replicated
//  solids' PressureP together; only_rebuild: the message only exists on steps that rebuild the buckets)
template <int SHIFT>
__global__ void k_wait(Ctl *ctl, unsigned long long epoch, const unsigned long long *flags, int nflags, int tag, int only_rebuild = 0,
                       const unsigned long long *flags2 = nullptr, int nflags2 = 0)
{
    if (only_rebuild && !ctl->rebuild) return;
    const int lane = threadIdx.x;
    const unsigned long long want = epoch;
    unsigned long long v = want << SHIFT;
    if (lane == 0) trace_mark(ctl, 100 + tag);
    if (lane < nflags + nflags2) {
        const unsigned long long *f = lane < nflags ? flags + lane : flags2 + (lane - nflags);
        const unsigned long long t0 = global_ns();
        for (;;) {
            v = ld_flag(f);
            if ((v >> SHIFT) >= want) break;
            if (global_ns() - t0 > kWaitTimeoutNs) { atomicOr(&ctl->err, kErrTimeout | (256 << tag)); break; }
            __nanosleep(64);
        }
    }
    __threadfence_system();
    if (SHIFT) {
        const unsigned any = __ballot_sync(0xffffffffu, lane < nflags && (v & 1ull));
        if (lane == 0 && any) ctl->need = 1;
    }
    __syncwarp();
    if (lane == 0) trace_mark(ctl, 200 + tag);
}
// this rank's rebuild vote (k_need's rule) into every rank's mailbox
__device__ __forceinline__ int local_need(const Ctl *ctl, const DecideArgs &a, const int *__restrict__ list_flags)
{
    const bool overflowed = list_flags && list_flags[0] != 0;
    return (ctl->force || !a.reuse_enabled || !ctl->skin_on || overflowed || __uint_as_float(ctl->maxdisp2) > a.half_skin2) ? 1 : 0;
}
__global__ void k_vote(Ctl *ctl, unsigned long long epoch, Peers peers, DecideArgs a, const int *__restrict__ list_flags)
{
    const int r = threadIdx.x;
    const int need = local_need(ctl, a, list_flags);
    if (r == 0) ctl->need = need;
    if (r < peers.nranks) st_flag(peers.vote[r] + peers.rank, (epoch << 1) | (unsigned long long)need);
}
// wait for every rank's vote, OR them, decide (k_decide's rule): one launch
__global__ void k_wait_decide(Ctl *ctl, unsigned long long epoch, const unsigned long long *flags, int nflags, DecideArgs a, int *__restrict__ list_flags)
{
    const int lane = threadIdx.x;
    unsigned long long v = epoch << 1;
    if (lane == 0) trace_mark(ctl, 100 + kWaitVote);
    if (lane < nflags) {
        const unsigned long long t0 = global_ns();
        for (;;) {
            v = ld_flag(flags + lane);
            if ((v >> 1) >= epoch) break;
            if (global_ns() - t0 > kWaitTimeoutNs) { atomicOr(&ctl->err, kErrTimeout | (256 << kWaitVote)); break; }
            __nanosleep(64);
        }
    }
    __threadfence_system();
    const unsigned any = __ballot_sync(0xffffffffu, lane < nflags && (v & 1ull));
    if (lane == 0) {
        if (any) ctl->need = 1;
        decide_body(ctl, a, list_flags);
        trace_mark(ctl, 200 + kWaitVote);
    }
}
// the last block of a pushing kernel publishes the message: count (optional), then the flag
__device__ __forceinline__ void push_complete(Ctl *ctl, unsigned long long epoch, int which, int *dst_cnt, int cnt, unsigned long long *dst_flag)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned done = atomicAdd(&ctl->push_done[which], 1u);
        if (done == gridDim.x - 1) {
            ctl->push_done[which] = 0;
            if (dst_cnt) *dst_cnt = cnt;
            __threadfence_system();
            st_flag(dst_flag, epoch);
            trace_mark(ctl, 400 + which);
        }
    }
}
// copy cnt * width doubles of a packed message into the neighbour's mailbox (coalesced stores over NVLink); both
// directions in one launch: blockIdx.y = 0 to the left neighbour, 1 to the right.  The count is clamped to the message
// capacity where it is published (cnt_p is written back).  only_rebuild: the message only exists on rebuild steps.
struct PushPair {
    const double *src[2];
    int *cnt[2];               // this context's counts (ctl->mig_cnt / ctl->halo_cnt)
    double *dst[2];            // the neighbours' mailbox buffers
    int *dst_cnt[2];
    unsigned long long *dst_flag[2];
};
__global__ void k_push(Ctl *ctl, unsigned long long epoch, int which0, PushPair pp, int width, int cap, int only_rebuild)
{
    if (only_rebuild && !ctl->rebuild) return;
    const int dir = blockIdx.y;
    const int cnt = min(*pp.cnt[dir], cap);
    const long long total = (long long)cnt * width;
    const double *__restrict__ src = pp.src[dir];
    double *__restrict__ dst = pp.dst[dir];
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) dst[k] = src[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned done = atomicAdd(&ctl->push_done[which0 + dir], 1u);
        if (done == gridDim.x - 1) {
            ctl->push_done[which0 + dir] = 0;
            *pp.cnt[dir] = cnt;
            *pp.dst_cnt[dir] = cnt;
            __threadfence_system();
            st_flag(pp.dst_flag[dir], epoch);
            trace_mark(ctl, 400 + which0 + dir);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K0: pre-step for every particle this context moves.  wall kinematics (t<0.2), periodic wrap, bucket key,
// and the displacement from the position the current candidate list was built on (the rebuild criterion).
// Solids take their state from the solid arrays (they are integrated there).  Ghosts are left alone: they
// die (rebuild) or are refreshed by their owner (reuse).
//   `only` != nullptr: the listed slots (the solids, handled once their sub-steps have finished);
//   skip_solids: everything but the solids.
__device__ __forceinline__ void pack_particle(double *dst, double x, double y, double z, double vx, double vy, double vz,
                                              int type, int id)
{
    dst[0] = x; dst[1] = y; dst[2] = z; dst[3] = vx; dst[4] = vy; dst[5] = vz;
    dst[6] = __longlong_as_double(((long long)real_type(type) << 32) | (long long)(unsigned)id);
}

__global__ void k_prestep(Ctl *ctl, Particles p, Solid sol, GridDesc g, WallMotion wm, int do_wrap,
                          const double *__restrict__ ancx, const double *__restrict__ ancy, const double *__restrict__ ancz,
                          const int *__restrict__ only, int only_count, int skip_solids)
{
    const int n = only ? only_count : ctl->n;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (t0 < n) {
        const int i = only ? only[t0] : t0;
        const int t = p.type[i];
        if (!(t & kGhost) && !(skip_solids && is_structure_type(t))) {
            double x = p.x[i], y = p.y[i], z = p.z[i];
            int s = -1;
            if (is_structure_type(t)) {
                s = p.id[i] - sol.sb;
                x = sol.x[s]; y = sol.y[s]; z = sol.z[s];
                p.vx[i] = sol.vx[s]; p.vy[i] = sol.vy[s]; p.vz[i] = sol.vz[s];
            } else if (wm.active && is_wall_type(t)) { // :3036-3060, operand order kept (bit-exact)
                const int tt = real_type(t);
                const double r0 = __dsub_rn(x, wm.center[tt][0]), r1 = __dsub_rn(y, wm.center[tt][1]),
                             r2 = __dsub_rn(z, wm.center[tt][2]);
                const double(*R)[3] = wm.R[tt];
                const double q0 = __dadd_rn(__dadd_rn(__dmul_rn(R[0][0], r0), __dmul_rn(R[0][1], r1)), __dmul_rn(R[0][2], r2));
                const double q1 = __dadd_rn(__dadd_rn(__dmul_rn(R[1][0], r0), __dmul_rn(R[1][1], r1)), __dmul_rn(R[1][2], r2));
                const double q2 = __dadd_rn(__dadd_rn(__dmul_rn(R[2][0], r0), __dmul_rn(R[2][1], r1)), __dmul_rn(R[2][2], r2));
                const double *w = wm.omega[tt], *V = wm.vel[tt];
                p.vx[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[1], q2), __dmul_rn(w[2], q1)), V[0]);
                p.vy[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[2], q0), __dmul_rn(w[0], q2)), V[1]);
                p.vz[i] = __dadd_rn(__dsub_rn(__dmul_rn(w[0], q1), __dmul_rn(w[1], q0)), V[2]);
                x = __dadd_rn(__dadd_rn(q0, wm.center[tt][0]), __dmul_rn(V[0], wm.dt));
                y = __dadd_rn(__dadd_rn(q1, wm.center[tt][1]), __dmul_rn(V[1], wm.dt));
                z = __dadd_rn(__dadd_rn(q2, wm.center[tt][2]), __dmul_rn(V[2], wm.dt));
            }
            if (do_wrap) { // :3330
                x = __dadd_rn(mod_exact(__dsub_rn(x, g.mn0g), g.W[0]), g.mn0g);
                y = __dadd_rn(mod_exact(__dsub_rn(y, g.mn[1]), g.W[1]), g.mn[1]);
                z = __dadd_rn(mod_exact(__dsub_rn(z, g.mn[2]), g.W[2]), g.mn[2]);
            }
            p.x[i] = x; p.y[i] = y; p.z[i] = z;
            if (s >= 0) {
                sol.x[s] = x; sol.y[s] = y; sol.z[s] = z;
                Rec u; // :2712, from the (wrapped) position the sub-steps start from
                u.a = minimg_exact(x, sol.x0[s], g.W[0]); u.b = minimg_exact(y, sol.y0[s], g.W[1]); u.c = minimg_exact(z, sol.z0[s], g.W[2]);
                u.d = 0.0;
                sol.u[s] = u;
            }
            p.key[i] = cell_key(g, x, y, z);
            // (a particle that wrapped through the box shows up a box width away: conservative, it forces a rebuild)
            const double ex = x - ancx[i], ey = y - ancy[i], ez = z - ancz[i];
            const double e2 = ex * ex + ey * ey + ez * ez;
            d2 = (e2 == e2) ? __double2float_ru(e2) : 3.0e38f; // NaN: rebuild (and the health flag reports it)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    // (one contended address: only a warp that raises the maximum issues the atomic)
    if ((threadIdx.x & 31) == 0 && d2 > __uint_as_float(*reinterpret_cast<volatile unsigned *>(&ctl->maxdisp2)))
        atomicMax(&ctl->maxdisp2, __float_as_uint(d2));
}

// The rebuild decision of this step, made on the device (one thread).  The list built with a skin d stays a
// superset of every cut-off set while no particle has moved further than d/2 from its build position.
//   * slab mode: `ctl->need` already holds the OR of all ranks' votes (k_wait<1>), so every rank decides alike;
//   * a list that overflowed (its particles need the bucket-walking fall-back) or was built without a skin is
//     rebuilt every step;
//   * the skin costs candidates (all the physics kernels traverse the longer lists), so it is only kept
//     while it pays: a list that served fewer than 2 steps switches the skin off for the following builds,
//     every 16th build tries it again.
// local vote (before the exchange of votes in slab mode)
__global__ void k_need(Ctl *ctl, DecideArgs a, const int *__restrict__ list_flags) { ctl->need = local_need(ctl, a, list_flags); }
__device__ __forceinline__ void decide_body(Ctl *ctl, const DecideArgs &a, int *__restrict__ list_flags)
{
    const int rebuild = ctl->need ? 1 : 0;
    ctl->rebuild = rebuild;
    ctl->maxdisp2 = 0u;
    ctl->mig_cnt[0] = 0; ctl->mig_cnt[1] = 0;
    if (rebuild) {
        ctl->last_age = ctl->age;
        ctl->age = 1;
        int skin = 0;
        if (a.reuse_enabled) {
            if (ctl->force || ctl->last_age >= 2 || ++ctl->probe >= 16) { skin = 1; ctl->probe = 0; }
        }
        ctl->skin_on = skin;
        ctl->filt2 = skin ? a.filt2_skin : a.filt2_plain;
        ctl->force = 0;
        ctl->halo_cnt[0] = 0; ctl->halo_cnt[1] = 0;
        ctl->ghost_cnt[0] = 0; ctl->ghost_cnt[1] = 0;
        ctl->n_own_sol = 0;
        ++ctl->builds;
        if (list_flags) list_flags[0] = 0;
    } else {
        ++ctl->age;
        ++ctl->reuses;
    }
}
__global__ void k_decide(Ctl *ctl, DecideArgs a, int *__restrict__ list_flags) { decide_body(ctl, a, list_flags); }

// K1 (rebuild steps): per-bucket count + arrival slot.  Slab mode: ghosts of the previous list die; fluid/wall
// particles whose column left the owned range are packed for the neighbouring slab (migration) and die here.
struct SlabSend {
    double *buf[2]; // local staging: [0] to the left neighbour, [1] to the right; 7 doubles per particle (AoS)
    int capacity;   // particles per buffer
};
__global__ void k_count(Ctl *ctl, Particles p, GridDesc g, int *__restrict__ cellCount, int *__restrict__ slot, SlabSend snd)
{
    if (!ctl->rebuild) return;
    const int n = ctl->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int t = p.type[i];
        int k;
        if (t & kGhost) k = g.ncells + 1; // last list's halo copy
        else {
            k = p.key[i];
            if (g.slab && !is_structure_type(t)) {
                if (k >= g.ncells) { atomicOr(&ctl->err, kErrLost); k = g.ncells + 1; } // moved further than one halo width: lost
                else {
                    const int col = key_column(g, k);
                    if (!column_owned(g, col)) { // migrate: hand the particle to the neighbouring slab
                        const int dir = (col < g.range) ? 0 : 1;
                        const int q = atomicAdd(&ctl->mig_cnt[dir], 1);
                        if (q < snd.capacity)
                            pack_particle(snd.buf[dir] + (size_t)kMsgDoubles * q, p.x[i], p.y[i], p.z[i], p.vx[i], p.vy[i], p.vz[i], t, p.id[i]);
                        else atomicOr(&ctl->err, kErrMsgFull);
                        k = g.ncells + 1;
                    }
                }
            }
        }
        p.key[i] = k;
        slot[i] = atomicAdd(&cellCount[k], 1);
    }
}
// slab mode, rebuild steps: append the particles received from a neighbour (migrants: ghost_flag 0; halo copies:
// ghost_flag kGhost, x shifted by +-W across the periodic seam so separations need no wrap in x).
// side 0 = from the left neighbour, 1 = from the right; cnt_in[2] are the received counts.
struct SidePair { const double *buf[2]; double xshift[2]; };
__global__ void k_unpack_particles(Ctl *ctl, SidePair sp, const int *cnt_in, int msg_cap, int cap, int ghost_flag, Particles p, GridDesc g,
                                   int *__restrict__ cellCount, int *__restrict__ slot, double *__restrict__ ancx, double *__restrict__ ancy,
                                   double *__restrict__ ancz)
{
    if (!ctl->rebuild) return;
    const int side = blockIdx.y;
    const int cl = min(max(cnt_in[0], 0), msg_cap), cr = min(max(cnt_in[1], 0), msg_cap);
    const double *__restrict__ buf = sp.buf[side];
    const double xshift = sp.xshift[side];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < (side ? cr : cl); q += gridDim.x * blockDim.x) {
        const int i = ctl->n + (side ? cl : 0) + q;
        if (i >= cap) { atomicOr(&ctl->err, kErrCapacity); return; }
        const double *m = buf + (size_t)kMsgDoubles * q;
        const double x = m[0] + xshift, y = m[1], z = m[2];
        const long long meta = __double_as_longlong(m[6]);
        p.x[i] = x; p.y[i] = y; p.z[i] = z; p.vx[i] = m[3]; p.vy[i] = m[4]; p.vz[i] = m[5];
        p.type[i] = (int)(meta >> 32) | ghost_flag;
        p.id[i] = (int)(meta & 0xffffffffLL);
        ancx[i] = x; ancy[i] = y; ancz[i] = z;
        int col;
        int k = cell_key(g, x, y, z, &col);
        const bool owned = column_owned(g, col);
        if (k >= g.ncells || (ghost_flag ? owned : !owned)) { atomicOr(&ctl->err, kErrArrival); k = g.ncells + 1; }
        p.key[i] = k;
        slot[i] = atomicAdd(&cellCount[k], 1);
    }
}
__global__ void k_advance_n(Ctl *ctl, const int *cnt_in, int msg_cap, int cap, int ghost)
{
    if (!ctl->rebuild) return;
    const int cl = min(max(cnt_in[0], 0), msg_cap), cr = min(max(cnt_in[1], 0), msg_cap);
    if (cnt_in[0] > msg_cap || cnt_in[1] > msg_cap) atomicOr(&ctl->err, kErrMsgFull);
    const int n = ctl->n;
    if (ghost) { ctl->ghost_base[0] = n; ctl->ghost_cnt[0] = cl; ctl->ghost_base[1] = n + cl; ctl->ghost_cnt[1] = cr; }
    ctl->n = min(n + cl + cr, cap);
}

// slab mode, rebuild steps: pack the owned fluid/wall particles within one halo width of the slab faces
__global__ void k_halo_pack(Ctl *ctl, Particles p, GridDesc g, SlabSend snd, int *__restrict__ haloSrc0, int *__restrict__ haloSrc1)
{
    if (!ctl->rebuild) return;
    const int n = ctl->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = p.key[i], t = p.type[i];
        if (k >= g.ncells || (t & kGhost) || is_structure_type(t)) continue;
        const int col = key_column(g, k);
        if (!column_owned(g, col)) continue;
        for (int dir = 0; dir < 2; ++dir) {
            const bool in = dir == 0 ? (col < 2 * g.range) : (col >= g.nx - 2 * g.range);
            if (!in) continue;
            const int q = atomicAdd(&ctl->halo_cnt[dir], 1);
            if (q < snd.capacity) {
                pack_particle(snd.buf[dir] + (size_t)kMsgDoubles * q, p.x[i], p.y[i], p.z[i], p.vx[i], p.vy[i], p.vz[i], t, p.id[i]);
                (dir == 0 ? haloSrc0 : haloSrc1)[q] = i;
            } else atomicOr(&ctl->err, kErrMsgFull);
        }
    }
}
// reuse steps: the SAME halo particles (sorted slots haloSlot, fixed when the list was built) with their current state,
// packed straight into the neighbours' mailboxes (blockIdx.y = direction) and completed by the flag: one launch
__global__ void k_halo_resend(Ctl *ctl, unsigned long long epoch, int which0, Particles p, const int *__restrict__ haloSlot0,
                              const int *__restrict__ haloSlot1, PushPair pp)
{
    if (ctl->rebuild) return;
    // (a particle's 7 doubles are gathered into shared memory, the block's tile leaves as one contiguous stream: NVLink moves
    //  128-byte packets about as fast as 8-byte ones)
    __shared__ double tile[256 * kMsgDoubles];
    const int dir = blockIdx.y;
    const int cnt = ctl->halo_cnt[dir];
    const int *__restrict__ hs = dir == 0 ? haloSlot0 : haloSlot1;
    double *__restrict__ dst = pp.dst[dir];
    for (int q0 = blockIdx.x * blockDim.x; q0 < cnt; q0 += gridDim.x * blockDim.x) {
        const int q = q0 + threadIdx.x;
        if (q < cnt) {
            const int i = hs[q];
            pack_particle(tile + (size_t)kMsgDoubles * threadIdx.x, p.x[i], p.y[i], p.z[i], p.vx[i], p.vy[i], p.vz[i], p.type[i], p.id[i]);
        }
        __syncthreads();
        const int nd = min((int)blockDim.x, cnt - q0) * kMsgDoubles;
        double *__restrict__ out = dst + (size_t)kMsgDoubles * q0;
        for (int k = threadIdx.x; k < nd; k += blockDim.x) out[k] = tile[k];
        __syncthreads();
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&ctl->push_done[which0 + dir], 1u) == gridDim.x - 1) {
            ctl->push_done[which0 + dir] = 0;
            *pp.dst_cnt[dir] = cnt;
            __threadfence_system();
            st_flag(pp.dst_flag[dir], epoch);
            trace_mark(ctl, 400 + which0 + dir);
        }
    }
}
// reuse steps: refresh the ghost slots with the state their owners sent (blockIdx.y = side)
__global__ void k_unpack_refresh(Ctl *ctl, SidePair sp, Particles p, const int *__restrict__ ghostSlot0, const int *__restrict__ ghostSlot1)
{
    if (ctl->rebuild) return;
    const int side = blockIdx.y;
    const int cnt = ctl->ghost_cnt[side];
    const double *__restrict__ buf = sp.buf[side];
    const int *__restrict__ gs = side == 0 ? ghostSlot0 : ghostSlot1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        const double *m = buf + (size_t)kMsgDoubles * q;
        const int i = gs[q];
        p.x[i] = m[0] + sp.xshift[side]; p.y[i] = m[1]; p.z[i] = m[2]; p.vx[i] = m[3]; p.vy[i] = m[4]; p.vz[i] = m[5];
    }
}
// after the permute of a rebuild step: the sorted slots of the halo particles sent / the ghosts received
__global__ void k_slab_slots(Ctl *ctl, const int *__restrict__ where, const int *__restrict__ haloSrc0, const int *__restrict__ haloSrc1,
                             int *__restrict__ haloSlot0, int *__restrict__ haloSlot1, int *__restrict__ ghostSlot0, int *__restrict__ ghostSlot1)
{
    if (!ctl->rebuild) return;
    const int m = max(max(ctl->halo_cnt[0], ctl->halo_cnt[1]), max(ctl->ghost_cnt[0], ctl->ghost_cnt[1]));
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m; q += gridDim.x * blockDim.x) {
        if (q < ctl->halo_cnt[0]) haloSlot0[q] = where[haloSrc0[q]];
        if (q < ctl->halo_cnt[1]) haloSlot1[q] = where[haloSrc1[q]];
        if (q < ctl->ghost_cnt[0]) ghostSlot0[q] = where[ctl->ghost_base[0] + q];
        if (q < ctl->ghost_cnt[1]) ghostSlot1[q] = where[ctl->ghost_base[1] + q];
    }
}
// second exchange: what pass 1 computed for the halo particles, in the order they were packed, straight into the
// neighbour's mailbox: PressureP (plane 0) and, with surface tension, PressureA and GravityCenter (planes 1..4);
// plane p of halo particle q at dst[p * msg_cap + q]
struct PassOneFields { const double *a[5]; int count; };
struct PassOneTargets { double *a[5]; int count; };
struct ScalarPush { double *dst[2]; unsigned long long *dst_flag[2]; };
__global__ void k_push_scalar(Ctl *ctl, unsigned long long epoch, int which0, const int *__restrict__ haloSlot0, const int *__restrict__ haloSlot1,
                              PassOneFields f, int msg_cap, ScalarPush sp)
{
    const int dir = blockIdx.y;
    const int cnt = ctl->halo_cnt[dir];
    const int *__restrict__ hs = dir == 0 ? haloSlot0 : haloSlot1;
    double *__restrict__ dst = sp.dst[dir];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        const int i = hs[q];
        for (int p = 0; p < f.count; ++p) dst[(size_t)p * msg_cap + q] = f.a[p][i];
    }
    push_complete(ctl, epoch, which0 + dir, nullptr, 0, sp.dst_flag[dir]);
}
__global__ void k_unpack_scalar(Ctl *ctl, const int *__restrict__ ghostSlot0, const int *__restrict__ ghostSlot1, const double *__restrict__ buf0,
                                const double *__restrict__ buf1, int msg_cap, PassOneTargets f, Rec *__restrict__ rb)
{
    const int side = blockIdx.y;
    const int cnt = ctl->ghost_cnt[side];
    const int *__restrict__ gs = side == 0 ? ghostSlot0 : ghostSlot1;
    const double *__restrict__ buf = side == 0 ? buf0 : buf1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        const int w = gs[q];
        for (int p = 0; p < f.count; ++p) f.a[p][w] = buf[(size_t)p * msg_cap + q];
        rb[w].c = buf[q]; // PressureP slot of the gather record
    }
}
// replicated solids: the slab that owns a solid particle (kSolidOwned: by its column when the list was built)
// computes its PressureP / its fluid-coupled velocity update and stores them into EVERY rank's mailbox.
__global__ void k_solid_owned_list(Ctl *ctl, Particles p, Solid sol, int *__restrict__ own_sol)
{
    if (!ctl->rebuild) return;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < sol.ns; s += gridDim.x * blockDim.x) {
        const int q = sol.slot[s];
        if (p.type[q] & kSolidOwned) own_sol[atomicAdd(&ctl->n_own_sol, 1)] = q;
    }
}
__global__ void k_solid_publish_P(Ctl *ctl, unsigned long long epoch, int which, Particles p, Solid sol, const int *__restrict__ own_sol, const double *__restrict__ P, Peers peers)
{
    const int cnt = ctl->n_own_sol;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        const int i = own_sol[q];
        const int s = p.id[i] - sol.sb;
        const double v = P[i];
        for (int r = 0; r < peers.nranks; ++r) peers.solP[r][s] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&ctl->push_done[which], 1u) == gridDim.x - 1) {
            ctl->push_done[which] = 0;
            __threadfence_system();
            for (int r = 0; r < peers.nranks; ++r) st_flag(peers.fsolP[r] + peers.rank, epoch);
        }
    }
}
// (with surface tension -- PA != nullptr -- the replicated solids also get what calculateDensityA / GravityCenter /
// PressureA give a structure particle: DensityA = 0, GravityCenter = 0 (:2149, :2183: i not structure), PressureA from nA = 0)
__global__ void k_solid_spread_P(Ctl *ctl, Particles p, GridDesc g, Solid sol, const double *__restrict__ solP, double *__restrict__ P,
                                 double *__restrict__ PA, double *__restrict__ gcx, double *__restrict__ gcy, double *__restrict__ gcz, Phys ph)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < sol.ns; s += gridDim.x * blockDim.x) {
        const int q = sol.slot[s];
        if (p.key[q] >= g.ncells || (p.type[q] & kGhost)) continue; // (parked outside this slab: never traversed)
        const double v = solP[s];
        P[q] = v;
        p.rb[q].c = v;
        if (PA) {
            double pa = ph.cofa[real_type(p.type[q])] * (0.0 - ph.n0a) / ph.l0; // :2219
            if (ph.n0a <= 0.0) pa = 0.0;
            PA[q] = pa; gcx[q] = 0.0; gcy[q] = 0.0; gcz[q] = 0.0;
        }
    }
}
// the owner's coupled velocities (pass 2 wrote them into its solid arrays) -> every rank's mailbox
__global__ void k_solid_publish_V(Ctl *ctl, unsigned long long epoch, int which, Particles p, Solid sol, const int *__restrict__ own_sol, Peers peers)
{
    const int cnt = ctl->n_own_sol;
    const size_t ns = sol.ns;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += gridDim.x * blockDim.x) {
        const int s = p.id[own_sol[q]] - sol.sb;
        const double vx = sol.vx[s], vy = sol.vy[s], vz = sol.vz[s];
        for (int r = 0; r < peers.nranks; ++r) {
            double *d = peers.solV[r];
            d[s] = vx; d[ns + s] = vy; d[2 * ns + s] = vz;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&ctl->push_done[which], 1u) == gridDim.x - 1) {
            ctl->push_done[which] = 0;
            __threadfence_system();
            for (int r = 0; r < peers.nranks; ++r) st_flag(peers.fsolV[r] + peers.rank, epoch);
        }
    }
}
// [s_lo, s_hi): the solids whose sub-steps this rank runs (they take the owners' coupled velocity); every rank clears the
// Force of the clamped solids it may have to report (updateElasticPosition does that inside the sub-steps, which only the
// rank that advances a solid runs; the clamp predicate is static)
__host__ __device__ inline bool solid_clamped(int module, double x0, double y0);
__global__ void k_solid_apply_update(Solid sol, const double *__restrict__ solV, int s_lo, int s_hi, int module)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sol.ns) return;
    const size_t ns = sol.ns;
    if (s >= s_lo && s < s_hi) { sol.vx[s] = solV[s]; sol.vy[s] = solV[ns + s]; sol.vz[s] = solV[2 * ns + s]; }
    else if (module != 0 && module != 3 && solid_clamped(module, sol.x0[s], sol.y0[s])) { sol.fx[s] = 0.0; sol.fy[s] = 0.0; sol.fz[s] = 0.0; }
}

// ------------------------------------------------------------------------------------------------
// K2: exclusive scan of the bucket counts (three small kernels; int32), rebuild steps only
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

// `gate` != nullptr: the kernel only works in rebuild steps (gate->rebuild)
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const Ctl *gate, const int *__restrict__ in, int n, int *__restrict__ blockSums)
{
    if (gate && !gate->rebuild) return;
    __shared__ int total;
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += in[base + k];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) k_scan_top(const Ctl *gate, int *__restrict__ blockSums, int nb)
{
    if (gate && !gate->rebuild) return;
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
// `clear`: zero the counts once they are scanned (the bucket counters are all-zero outside a rebuild)
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const Ctl *gate, int *__restrict__ in, int n, const int *__restrict__ blockPrefix,
                                                             int *__restrict__ out /* n+1 */, int clear)
{
    if (gate && !gate->rebuild) return;
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
        if (base + k < n) { out[base + k] = ex; if (clear) in[base + k] = 0; }
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

// K3: provisional bucket order (arrival order inside a bucket is arbitrary -> fixed up in K4)
__global__ void k_scatter_index(const Ctl *ctl, const int *__restrict__ key, const int *__restrict__ slot,
                                const int *__restrict__ cellStart, int *__restrict__ tmpIdx)
{
    if (!ctl->rebuild) return;
    const int n = ctl->n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) tmpIdx[cellStart[key[i]] + slot[i]] = i;
}

// K4: permute the SoA into bucket order and write the gather records.  Inside a bucket particles are ordered by
// original id, which makes the layout (and therefore every floating-point sum) independent of atomic arrival
// order: the thread of an arrival ranks its own id among the bucket's occupants (O(m) per thread).
// Reuse steps keep the order (identity permutation): only the state and the records move.
__global__ void k_permute(const Ctl *ctl, Particles src, Particles dst, const int *__restrict__ cellStart,
                          const int *__restrict__ tmpIdx, GridDesc g, int *__restrict__ where, int *__restrict__ solid_slot,
                          int solid_base, double *__restrict__ ancx, double *__restrict__ ancy, double *__restrict__ ancz)
{
    const int rebuild = ctl->rebuild;
    const int q0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= (rebuild ? cellStart[g.ncells + 1] : ctl->n)) return; // the dead bucket (last) is dropped
    const int s = rebuild ? tmpIdx[q0] : q0;
    const int k = src.key[s];
    int q = q0;
    if (rebuild && k < g.ncells) { // (the parked bucket is never traversed: arrival order will do)
        const int b = cellStart[k], e = cellStart[k + 1];
        if (e - b > 1) {
            const int ids = src.id[s];
            int rank = 0;
            for (int c = b; c < e; ++c) rank += (src.id[tmpIdx[c]] < ids);
            q = b + rank;
        }
    }
    const double x = src.x[s], y = src.y[s], z = src.z[s];
    int t = src.type[s];
    const bool solid = is_structure_type(t) && !(t & kGhost);
    if (rebuild) {
        where[s] = q;
        if (solid_slot && solid) solid_slot[src.id[s] - solid_base] = q; // sorted slot of each solid
        if (g.slab && solid) { // the slab that owns the solid's column evaluates it until the next rebuild
            const bool own = k < g.ncells && column_owned(g, key_column(g, k));
            t = (t & ~kSolidOwned) | (own ? kSolidOwned : 0);
        }
        ancx[q] = x; ancy[q] = y; ancz[q] = z; // the positions this list is built on
    }
    dst.x[q] = x; dst.y[q] = y; dst.z[q] = z;
    dst.type[q] = t; dst.id[q] = src.id[s]; dst.key[q] = k;
    const double icw = 1.0 / g.cellw;
    // filter coordinates in bucket units, consistent with the (periodically wrapped) bucket index: a
    // position that has not been wrapped into the box yet (input of the very first bucket build, which
    // the reference does before its first calculatePeriodicBoundary) is filtered at its image in the box
    double fx = (x - g.mn[0]) * icw, fy = (y - g.mn[1]) * icw, fz = (z - g.mn[2]) * icw;
    double xr = x; // x as a NEIGHBOUR's thread sees it (the gather record)
    if (!g.slab) fx = fx < 0.0 ? fx + g.nx : (fx >= g.nx ? fx - g.nx : fx);
    else if (fx < 0.0) { fx += g.nxg; xr = x + g.W[0]; }          // slab mode: a replicated solid keeps its global x; seen
    else if (fx >= (double)g.nxg) { fx -= g.nxg; xr = x - g.W[0]; } // through the periodic seam it sits one box width away
                                                                   // (ghost fluid arrives already shifted)
    if (rebuild) { // (only the filter reads pf)
        fy = fy < 0.0 ? fy + g.ny : (fy >= g.ny ? fy - g.ny : fy);
        fz = fz < 0.0 ? fz + g.nz : (fz >= g.nz ? fz - g.nz : fz);
        float *pfw = reinterpret_cast<float *>(dst.pf + (q >> 1)) + (q & 1); // x0 x1 y0 y1 z0 z1
        pfw[0] = (float)fx; pfw[2] = (float)fy; pfw[4] = (float)fz;
    }
    const double vx = src.vx[s], vy = src.vy[s], vz = src.vz[s];
    dst.vx[q] = vx; dst.vy[q] = vy; dst.vz[q] = vz;
    Rec ra, rb;
    ra.a = xr; ra.b = y; ra.c = z; ra.d = vx;
    rb.a = vy; rb.b = vz; rb.c = 0.0; rb.d = __longlong_as_double((long long)real_type(t));
    dst.ra[q] = ra;
    dst.rb[q] = rb;
}
// after the permute of a rebuild step: the live slots (everything but the dead bucket)
__global__ void k_set_n(Ctl *ctl, const int *__restrict__ cellStart, int ncells)
{
    if (ctl->rebuild) ctl->n = cellStart[ncells + 1];
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
    // slab mode: rows are produced by the slab that owns the particle's column (never for ghosts)
    if (key[i] >= g.ncells || (type[i] & kGhost) || !column_owned(g, key_column(g, key[i]))) return;
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

// a bucket structure over an arbitrary particle set, independent of the stepping state: serves the exact
// neighbour lists above (debug / VTK output) and the initial structure lists (once)
__global__ void k_dbg_keycount(int n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z, GridDesc g,
                               int *__restrict__ key, int *__restrict__ cellCount, int *__restrict__ slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = cell_key(g, x[i], y[i], z[i]);
    key[i] = k;
    slot[i] = atomicAdd(&cellCount[k], 1);
}
__global__ void k_dbg_gather(int n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                             const int *__restrict__ type, const int *__restrict__ id, const int *__restrict__ key,
                             const int *__restrict__ slot, const int *__restrict__ cellStart, double *__restrict__ ox, double *__restrict__ oy,
                             double *__restrict__ oz, int *__restrict__ otype, int *__restrict__ oid, int *__restrict__ okey)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = cellStart[key[i]] + slot[i];
    ox[q] = x[i]; oy[q] = y[i]; oz[q] = z[i];
    otype[q] = type[i]; oid[q] = id[i]; okey[q] = key[i];
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
// PACKED: the static pair data comes from the tuple dictionary (see Solid).
// [s_lo, s_hi): the solids this context advances (all of them on a single context, a share on a slab).
//
// The sub-step kernels are latency bound (~10^5 threads, two dependent loads per pair: list entry -> gather): the
// list entries of the NEXT batch of kSolidBatch pairs are fetched while the current batch is processed, and the
// gathers of a few pairs are issued together.  The sums keep the reference's serial order.
// (Measured on B200, 113k solids, per step of 5 sub-steps: plain loop 0.74 ms; this form 0.61 ms; four lanes per
// solid -- one per Cartesian component -- 1.36 ms: the per-thread registers stay, so 6x fewer solids are in flight.)
#ifndef MPHX_SOLID_BATCH
#define MPHX_SOLID_BATCH 8
#endif
#ifndef MPHX_S1_MINB
#define MPHX_S1_MINB 5
#endif
#ifndef MPHX_S2_MINB
#define MPHX_S2_MINB 6
#endif
#ifndef MPHX_S2_G
#define MPHX_S2_G 2 // pairs whose gathers are issued together in pass 2 (9 + 4 doubles each)
#endif
constexpr int kSolidBatch = MPHX_SOLID_BATCH;
__device__ __forceinline__ Rec ld_rec_ro(const Rec *p)
{
    Rec r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
// Split sub-steps (slab mode): rank r advances the solids [s_lo, s_hi) and stores what other ranks' rows read -- the
// stress P after pass 1, the displacement u after pass 2 -- straight into their solid arrays over NVLink (Solid::pmask:
// the ranks whose rows reference s), from the kernel that computes it.  The last block of a launch then raises this rank's
// phase counter in every rank's mailbox; the next kernel of every rank starts by waiting for all counters (ring_wait).
// The wait for the previous phase of every rank, inside the kernel that needs it: the kernel is launched (and its blocks
// resident) while the slower ranks still compute, so the launch latency hides behind the skew between the ranks, and a phase
// costs one launch instead of two.  All threads of the block call; the data the phase reads is only loaded afterwards.
__device__ __forceinline__ void ring_wait(Ctl *ctl, const SolidRing &ring)
{
    if (ring.wait_seq == 0) return;
    const bool marks = blockIdx.x == 0 && threadIdx.x == 0;
    if (marks) trace_mark(ctl, 100 + kWaitSub);
    if ((int)threadIdx.x < ring.nranks) {
        const unsigned long long *f = (const unsigned long long *)(ring.base[ring.rank] + ring.off_flag) + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_flag(f) < ring.wait_seq) {
            if (global_ns() - t0 > kWaitTimeoutNs) { atomicOr(&ctl->err, kErrTimeout | (256 << kWaitSub)); break; }
            __nanosleep(32);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (marks) trace_mark(ctl, 200 + kWaitSub);
}
// (the threads that stored into peer memory have fenced after their stores: solid_pass*_finish)
__device__ __forceinline__ void ring_complete(Ctl *ctl, int which, const SolidRing &ring)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&ctl->sub_done[which], 1u) == gridDim.x - 1) {
            ctl->sub_done[which] = 0;
            trace_mark(ctl, 320 + which);
            __threadfence_system();
            for (int r = 0; r < ring.nranks; ++r) st_flag((unsigned long long *)(ring.base[r] + ring.off_flag) + ring.rank, ring.seq);
            trace_mark(ctl, 300 + which);
        }
    }
}
template <int DIMS>
__device__ __forceinline__ void solid_pass1_finish(const Solid &so, const int s, const double (&G)[3][3], double (&Pv)[9]);
template <int DIMS, bool PACKED, bool DEEP>
__device__ __forceinline__ void solid_pass1_row(const Solid &so, const int s, double (&Pv)[9])
{
    constexpr int GI = DEEP ? 8 : 4; // pairs whose gathers are in flight together
    using namespace ex;
    const int ns = so.ns;
    const Rec uo = so.u[s];
    const double ui[3] = {uo.a, uo.b, DIMS == 3 ? uo.c : 0.0};
    double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int len = so.len[s];
    int jn[kSolidBatch];
    unsigned tn[kSolidBatch];
    auto fetch = [&](int kk0) {
#pragma unroll
        for (int u = 0; u < kSolidBatch; ++u) {
            const int kk = kk0 + u < len ? kk0 + u : (len > 0 ? len - 1 : 0);
            const size_t k = (size_t)kk * ns + s;
            jn[u] = len > 0 ? __ldg(&so.enbr[k]) : s;
            tn[u] = (PACKED && len > 0) ? (unsigned)__ldg(&so.tix[k]) : 0u;
        }
    };
    fetch(0);
    for (int kk0 = 0; kk0 < len; kk0 += kSolidBatch) {
        int jc[kSolidBatch];
        unsigned tc[kSolidBatch];
#pragma unroll
        for (int u = 0; u < kSolidBatch; ++u) { jc[u] = jn[u]; tc[u] = tn[u]; }
        if (kk0 + kSolidBatch < len) fetch(kk0 + kSolidBatch);
#pragma unroll
        for (int h = 0; h < kSolidBatch; h += GI) {
            Rec tt[GI], un[GI];
#pragma unroll
            for (int u = 0; u < GI; ++u) {
                const size_t k = (size_t)(kk0 + h + u < len ? kk0 + h + u : 0) * ns + s;
                if (PACKED) tt[u] = ld_rec_ro(so.ttab + tc[h + u]);
                else {
                    tt[u].a = __ldg(&so.d0x[k]); tt[u].b = __ldg(&so.d0y[k]); tt[u].c = DIMS == 3 ? __ldg(&so.d0z[k]) : 0.0;
                    tt[u].d = __ldg(&so.w[k]);
                }
                un[u] = so.u[jc[h + u]];
            }
#pragma unroll
            for (int u = 0; u < GI; ++u) {
                if (kk0 + h + u >= len) { if (DEEP) continue; else break; } // (DEEP: straight-line code, the tail is predicated)
                const double d0[3] = {tt[u].a, tt[u].b, DIMS == 3 ? tt[u].c : 0.0};
                const double w = tt[u].d;
                const double uj[3] = {un[u].a, un[u].b, DIMS == 3 ? un[u].c : 0.0};
                double d[3];
                for (int a = 0; a < DIMS; ++a) d[a] = add(d0[a], sub(uj[a], ui[a])); // :2716
                for (int a = 0; a < DIMS; ++a)
                    for (int b = 0; b < DIMS; ++b) G[a][b] = add(G[a][b], mul(mul(w, d[a]), d0[b])); // :2726
            }
        }
    }
    solid_pass1_finish<DIMS>(so, s, G, Pv);
}
// F = G L^-1 (:2743), E (:2780), S (:2804), P = F S L^-1 (:2847) of solid s from its accumulated G; P also comes back in Pv
template <int DIMS>
__device__ __forceinline__ void solid_pass1_finish(const Solid &so, const int s, const double (&G)[3][3], double (&Pv)[9])
{
    using namespace ex;
    const int ns = so.ns;
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
    double *Pk = so.PkA + 9 * (size_t)s;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            if (a >= DIMS || b >= DIMS) { Pk[3 * a + b] = 0.0; continue; }
            double sum = 0.0;
            for (int k = 0; k < DIMS; ++k)
                for (int l = 0; l < DIMS; ++l) sum = add(sum, mul(mul(F[a][k], S[k][l]), L[l][b])); // :2847
            Pk[3 * a + b] = sum;
            MPHX_T(so.Fm, a, b, s, ns) = F[a][b];
            MPHX_T(so.E, a, b, s, ns) = E[a][b];
            MPHX_T(so.S, a, b, s, ns) = S[a][b];
        }
#pragma unroll
    for (int e = 0; e < 9; ++e) Pv[e] = Pk[e];
}
// P of s -> the ranks whose pass 2 gathers it: scattered (a thread stores its own nine doubles) ...
__device__ __forceinline__ void ring_push_P(const Solid &so, const int s, const double (&Pv)[9], const SolidRing &ring)
{
    for (unsigned m = so.pmask[s]; m; m &= m - 1) {
        double *d = (double *)(ring.base[__ffs(m) - 1] + ring.off_pk) + 9 * (size_t)s;
#pragma unroll
        for (int e = 0; e < 9; ++e) d[e] = Pv[e];
    }
    __threadfence_system();
}
// ... or by the warp (one thread per solid, consecutive solids): the warp's 32 x 9 doubles are contiguous in PkA, so they
// go through shared memory and leave as 256-byte rows to every rank that needs any of them (all lanes must call)
__device__ __forceinline__ void ring_push_P_warp(const Solid &so, const int s, const bool active, const double (&Pv)[9], int s_hi,
                                                 const SolidRing &ring, double *stage /* [32 * 9] of this warp */)
{
    const int lane = threadIdx.x & 31;
    const unsigned um = __reduce_or_sync(0xffffffffu, active ? (unsigned)so.pmask[s] : 0u);
    if (um == 0u) return;
    if (active) {
#pragma unroll
        for (int e = 0; e < 9; ++e) stage[lane * 9 + e] = Pv[e];
    }
    __syncwarp();
    const int s0 = s - lane;
    const int nd = min(32, s_hi - s0) * 9;
    for (unsigned m = um; m; m &= m - 1) {
        double *d = (double *)(ring.base[__ffs(m) - 1] + ring.off_pk) + 9 * (size_t)s0;
        for (int k = lane; k < nd; k += 32) d[k] = stage[k];
    }
    __threadfence_system();
}
// (RING: a rank's share is ~ns/nranks solids, about one block per SM, and its kernels run beside the fluid's share of pass 2
// on a high-priority stream: blocks of kRingBlock threads with <= 128 registers need no more of an SM than ONE retiring
// pass-2 block frees (80 registers x 128 threads) -- a 218-register block of 128 threads waited for three to retire at once,
// which the scheduler does not arrange: the sub-steps stretched over the whole of pass 2)
// DEEP: eight pairs' gathers in flight per thread, one warp per block, up to 255 registers -- for a latency chain at about
// one warp per SM (a rank's share) registers are free and only memory-level parallelism shortens the chain.
constexpr int kRingBlock = 32;
template <int DIMS, bool PACKED, bool RING, bool DEEP>
__global__ void __launch_bounds__(DEEP ? kRingBlock : 128, DEEP ? 8 : MPHX_S1_MINB)
k_solid_pass1(Ctl *ctl, Solid so, int s_lo, int s_hi, SolidRing ring)
{
    __shared__ double stage[RING ? (DEEP ? kRingBlock : 128) * 9 : 1];
    const int s = s_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (RING) ring_wait(ctl, ring);
    if (RING && threadIdx.x == 0 && blockIdx.x == 0) trace_mark(ctl, 310);
    double Pv[9];
    if (s < s_hi) solid_pass1_row<DIMS, PACKED, DEEP>(so, s, Pv);
    if (RING) {
        ring_push_P_warp(so, s, s < s_hi, Pv, s_hi, ring, stage + (threadIdx.x / 32) * 32 * 9);
        ring_complete(ctl, 0, ring);
    }
}

// the clamp variants of updateElasticPosition (MPHX_MODULE_*: Bar :1919, DAM :1968, Turek_Hron :1944, Rolling1 :1992,
// Hydroelastic :2016, Rolling2 :2040)
__host__ __device__ inline bool solid_clamped(int module, double x0, double y0)
{
    switch (module) {
    case 1: return x0 < 0.001;
    case 2: return y0 < 0.002;
    case 3: return x0 < 0.205;
    case 4: return y0 < 0.003;
    case 5: return x0 < 0.01 || x0 > 1.99;
    case 6: return y0 > 0.3420;
    default: return false;
    }
}

// K8 "solid pass 2": the reference scatters  v_i += w P_i x0_ij /rho_i dt,  v_j -= (same)/rho_j dt
// serially / with atomics (:2855-2887); here every particle GATHERS, in the reference's serial
// order, the terms of rows j<s that list s, then its own row, then rows j>s (transposed list), so
// the result is deterministic, atomic-free and equal to the reference's CPU bits.
// Then updateElasticPosition (:1916-2081) incl. the clamp modules and quirk Q1.
template <bool RING>
__device__ __forceinline__ void solid_pass2_finish(const Solid &so, const int s, double (&v)[3], double W0, double W1, double W2, double edt,
                                                   int module, int double_update, const SolidRing &ring);
template <int DIMS, bool PACKED, bool RING, bool DEEP>
__device__ __forceinline__ void solid_pass2_row(const Solid &so, const int s, double W0, double W1, double W2, double edt, int module,
                                                int double_update, const double *__restrict__ inv_density, const SolidRing &ring)
{
    constexpr int GI = DEEP ? 8 : MPHX_S2_G; // pairs whose gathers are in flight together
    using namespace ex;
    const int ns = so.ns;
    const double ir = inv_density[so.type[s]];
    double v[3] = {so.vx[s], so.vy[s], so.vz[s]};
    // one row (own: sign +, P of s itself; transposed: sign -, P of the listing row j), pipelined like pass 1
    auto row = [&](const int *__restrict__ nbr, const unsigned short *__restrict__ tix, const double *__restrict__ ax, const double *__restrict__ ay,
                   const double *__restrict__ az, const double *__restrict__ aw, int kb, int ke, bool own, const double (&Pi)[3][3]) {
        if (ke - kb <= 0) return;
        int jn[kSolidBatch];
        unsigned tn[kSolidBatch];
        auto fetch = [&](int kk0) {
#pragma unroll
            for (int u = 0; u < kSolidBatch; ++u) {
                const int kk = kk0 + u < ke ? kk0 + u : ke - 1;
                const size_t k = (size_t)kk * ns + s;
                jn[u] = own ? s : __ldg(&nbr[k]);
                tn[u] = PACKED ? (unsigned)__ldg(&tix[k]) : 0u;
            }
        };
        fetch(kb);
        for (int kk0 = kb; kk0 < ke; kk0 += kSolidBatch) {
            int jc[kSolidBatch];
            unsigned tc[kSolidBatch];
#pragma unroll
            for (int u = 0; u < kSolidBatch; ++u) { jc[u] = jn[u]; tc[u] = tn[u]; }
            if (kk0 + kSolidBatch < ke) fetch(kk0 + kSolidBatch);
#pragma unroll
            for (int h = 0; h < kSolidBatch; h += GI) {
                Rec tt[GI];
                double Pj[GI][9];
#pragma unroll
                for (int u = 0; u < GI; ++u) {
                    const size_t k = (size_t)(kk0 + h + u < ke ? kk0 + h + u : kb) * ns + s;
                    if (PACKED) tt[u] = ld_rec_ro(so.ttab + tc[h + u]);
                    else {
                        tt[u].a = __ldg(&ax[k]); tt[u].b = __ldg(&ay[k]); tt[u].c = DIMS == 3 ? __ldg(&az[k]) : 0.0;
                        tt[u].d = __ldg(&aw[k]);
                    }
                    if (!own) {
                        const double *pp = so.PkA + 9 * (size_t)jc[h + u];
#pragma unroll
                        for (int e = 0; e < 9; ++e) Pj[u][e] = (e / 3 < DIMS && e % 3 < DIMS) ? pp[e] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < GI; ++u) {
                    if (kk0 + h + u >= ke) { if (DEEP) continue; else break; }
                    const double d0[3] = {tt[u].a, tt[u].b, DIMS == 3 ? tt[u].c : 0.0};
                    const double w = tt[u].d;
                    for (int a = 0; a < DIMS; ++a) {
                        double f = 0.0;
                        for (int b = 0; b < DIMS; ++b) f = add(f, mul(own ? Pi[a][b] : Pj[u][3 * a + b], d0[b]));
                        f = mul(f, w);
                        if (own) v[a] = add(v[a], mul(mul(ir, f), edt)); // :2883  v_s += invRho_s * (w P_s x0_sj) * dt
                        else v[a] = sub(v[a], mul(mul(ir, f), edt));     // :2885  row j lists s: v_s -= invRho_s * (w P_j x0_js) * dt
                    }
                }
            }
        }
    };
    const int rlen = so.rlen[s], rsplit = so.rsplit[s]; // transposed entries [0, rsplit) are rows j < s
    double Pi[3][3];
    {
        const double *Ps = so.PkA + 9 * (size_t)s;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) Pi[a][b] = (a < DIMS && b < DIMS) ? Ps[3 * a + b] : 0.0;
    }
    row(so.ernbr, so.rtix, so.rd0x, so.rd0y, so.rd0z, so.rw, 0, rsplit, false, Pi);
    row(so.enbr, so.tix, so.d0x, so.d0y, so.d0z, so.w, 0, so.len[s], true, Pi);
    row(so.ernbr, so.rtix, so.rd0x, so.rd0y, so.rd0z, so.rw, rsplit, rlen, false, Pi);
    solid_pass2_finish<RING>(so, s, v, W0, W1, W2, edt, module, double_update, ring);
}
// updateElasticPosition (:1916-2081) of solid s with its new velocity v, incl. the clamp modules and quirk Q1
template <bool RING>
__device__ __forceinline__ void solid_pass2_finish(const Solid &so, const int s, double (&v)[3], double W0, double W1, double W2, double edt,
                                                   int module, int double_update, const SolidRing &ring)
{
    using namespace ex;
    const double xi0 = so.x0[s], yi0 = so.y0[s], zi0 = so.z0[s];
    double x[3] = {so.x[s], so.y[s], so.z[s]};
    // Acceleration of solids is 0 (:2892): v += 0*dt leaves v unchanged
    if (module != 0) {
        if (solid_clamped(module, xi0, yi0)) {
            x[0] = xi0; x[1] = yi0; x[2] = zi0;
            v[0] = v[1] = v[2] = 0.0;
            if (module != 3) { so.fx[s] = 0.0; so.fy[s] = 0.0; so.fz[s] = 0.0; } // (Turek_Hron leaves Force alone, :1944-1953)
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
    Rec u;
    u.a = minimg_exact(x[0], xi0, W0); u.b = minimg_exact(x[1], yi0, W1); u.c = minimg_exact(x[2], zi0, W2); u.d = 0.0;
    so.u[s] = u;
    if (RING) {
        for (unsigned m = so.pmask[s]; m; m &= m - 1) // u of s -> the ranks whose pass 1 gathers it
            ((Rec *)(ring.base[__ffs(m) - 1] + ring.off_u))[s] = u;
        if (ring.last) // the step's final state of s -> every rank (the replicated solids of the fluid passes)
            for (int r = 0; r < ring.nranks; ++r) {
                if (r == ring.rank) continue;
                char *b = ring.base[r];
                ((double *)(b + ring.off_xv[0]))[s] = x[0]; ((double *)(b + ring.off_xv[1]))[s] = x[1]; ((double *)(b + ring.off_xv[2]))[s] = x[2];
                ((double *)(b + ring.off_xv[3]))[s] = v[0]; ((double *)(b + ring.off_xv[4]))[s] = v[1]; ((double *)(b + ring.off_xv[5]))[s] = v[2];
            }
        __threadfence_system();
    }
}
template <int DIMS, bool PACKED, bool RING, bool DEEP>
__global__ void __launch_bounds__(DEEP ? kRingBlock : 128, DEEP ? 8 : MPHX_S2_MINB)
k_solid_pass2(Ctl *ctl, Solid so, int s_lo, int s_hi, double W0, double W1, double W2, double edt, int module, int double_update,
              const double *__restrict__ inv_density, SolidRing ring)
{
    const int s = s_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (RING) ring_wait(ctl, ring);
    if (RING && threadIdx.x == 0 && blockIdx.x == 0) trace_mark(ctl, 311);
    if (s < s_hi) solid_pass2_row<DIMS, PACKED, RING, DEEP>(so, s, W0, W1, W2, edt, module, double_update, inv_density, ring);
    if (RING) ring_complete(ctl, 1, ring);
}

// ---- the sub-step kernels with a TEAM of lanes per solid (dictionary-packed solids) ------------------------------
// One thread per solid walks ~80 (pass 1) / ~160 (pass 2) pairs with a dependent load per pair: ~60 us per launch whatever
// the number of solids (measured: 113k solids on one GPU 58 us, a 28k share 72 us) -- a latency chain, 84 % long-scoreboard
// stalls.  The pair TERMS are independent; only their SUM has the reference's serial order.  So a team of kTeam lanes
// computes the terms of a solid in parallel (lane l takes pairs l, l + kTeam, ...: no loop-carried dependence, loads of
// several pairs in flight per lane, kTeam times more lanes in flight) into shared memory, then one lane per tensor / vector
// component adds them up in the serial order, and one thread per solid of the block does the small epilogue.  Same
// operations on the same operands in the same order as the one-thread kernels: the same bits.
// The lists are read in CSR order (Solid::off/nbr, roff/rnbr, ctix/crtix): a team's lanes read consecutive entries.
constexpr int kTeam = 16;
constexpr int kTeamBlock = 128; // 8 solids per block; a block needs no more registers than one retiring pass-2 block frees
// A chunk = kTeam consecutive pairs of a row (one per lane).  The terms of a chunk go through a double-buffered
// [kTeam][components] tile of shared memory; the gathers of the next chunk and the list entries of the one after are in
// flight while the current chunk is computed and summed.
template <int DIMS, bool RING>
__global__ void __launch_bounds__(kTeamBlock, 6) k_solid_pass1_team(Ctl *ctl, Solid so, int s_lo, int s_hi, SolidRing ring)
{
    using namespace ex;
    constexpr int NE = DIMS * DIMS, TEAMS = kTeamBlock / kTeam;
    __shared__ double Ts[TEAMS][2][kTeam][NE];
    __shared__ double Gs[TEAMS][NE];
    const int team = threadIdx.x / kTeam, lane = threadIdx.x % kTeam;
    const int s = s_lo + blockIdx.x * TEAMS + team;
    if (RING) ring_wait(ctl, ring);
    if (s < s_hi) { // (a team is half a warp: both halves take the same number of trips only by chance, hence the masks)
        const unsigned mask = 0xffffu << (16 * ((threadIdx.x / kTeam) & 1));
        const int len = so.len[s], base = so.off[s];
        const Rec uo = so.u[s];
        const double ui[3] = {uo.a, uo.b, DIMS == 3 ? uo.c : 0.0};
        double g = 0.0;
        // pipeline: (j, ti) of chunk c + 2, (t, un) of chunk c + 1
        int jn = 0; unsigned tin = 0;
        Rec tn{}, un{};
        if (lane < len) { jn = __ldg(&so.nbr[base + lane]); tin = __ldg(&so.ctix[base + lane]); }
        if (lane < len) { tn = ld_rec_ro(so.ttab + tin); un = so.u[jn]; }
        if (kTeam + lane < len) { jn = __ldg(&so.nbr[base + kTeam + lane]); tin = __ldg(&so.ctix[base + kTeam + lane]); }
        int buf = 0;
        for (int c0 = 0; c0 < len; c0 += kTeam, buf ^= 1) {
            const Rec t = tn, uj = un;
            if (c0 + kTeam + lane < len) { tn = ld_rec_ro(so.ttab + tin); un = so.u[jn]; }
            if (c0 + 2 * kTeam + lane < len) { jn = __ldg(&so.nbr[base + c0 + 2 * kTeam + lane]); tin = __ldg(&so.ctix[base + c0 + 2 * kTeam + lane]); }
            if (c0 + lane < len) {
                const double d0[3] = {t.a, t.b, DIMS == 3 ? t.c : 0.0};
                const double ujv[3] = {uj.a, uj.b, DIMS == 3 ? uj.c : 0.0};
                double d[3];
                for (int a = 0; a < DIMS; ++a) d[a] = add(d0[a], sub(ujv[a], ui[a])); // :2716
                for (int a = 0; a < DIMS; ++a)
                    for (int b2 = 0; b2 < DIMS; ++b2) Ts[team][buf][lane][a * DIMS + b2] = mul(mul(t.d, d[a]), d0[b2]); // the term of :2726
            }
            __syncwarp(mask);
            if (lane < NE) { // :2726, serial order (all of the chunk's terms are read first, then the dependent adds)
                const int m = len - c0;
                double tv[kTeam];
#pragma unroll
                for (int p = 0; p < kTeam; ++p) tv[p] = Ts[team][buf][p][lane];
#pragma unroll
                for (int p = 0; p < kTeam; ++p) g = p < m ? add(g, tv[p]) : g;
            }
        }
        if (lane < NE) Gs[team][lane] = g;
    }
    __syncthreads();
    if (threadIdx.x < TEAMS) {
        const int s2 = s_lo + blockIdx.x * TEAMS + threadIdx.x;
        if (s2 < s_hi) {
            double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int a = 0; a < DIMS; ++a)
                for (int b2 = 0; b2 < DIMS; ++b2) G[a][b2] = Gs[threadIdx.x][a * DIMS + b2];
            double Pv[9];
            solid_pass1_finish<DIMS>(so, s2, G, Pv);
            if (RING) ring_push_P(so, s2, Pv, ring);
        }
    }
    if (RING) ring_complete(ctl, 0, ring);
}
template <int DIMS, bool RING>
__global__ void __launch_bounds__(kTeamBlock, 6)
k_solid_pass2_team(Ctl *ctl, Solid so, int s_lo, int s_hi, double W0, double W1, double W2, double edt, int module, int double_update,
                   const double *__restrict__ inv_density, SolidRing ring)
{
    using namespace ex;
    constexpr int TEAMS = kTeamBlock / kTeam;
    __shared__ double Ts[TEAMS][2][kTeam][DIMS];
    __shared__ double Vs[TEAMS][3];
    const int team = threadIdx.x / kTeam, lane = threadIdx.x % kTeam;
    const int s = s_lo + blockIdx.x * TEAMS + team;
    if (RING) ring_wait(ctl, ring);
    if (s < s_hi) {
        const unsigned mask = 0xffffu << (16 * ((threadIdx.x / kTeam) & 1));
        // the reference's serial order (:2855-2887): rows j < s that list s, the own row, rows j > s
        const int len = so.len[s], rlen = so.rlen[s], rsplit = so.rsplit[s];
        const int base = so.off[s], rbase = so.roff[s];
        const int total = len + rlen;
        const double ir = inv_density[so.type[s]];
        double v = lane == 0 ? so.vx[s] : lane == 1 ? so.vy[s] : lane == 2 ? so.vz[s] : 0.0;
        // entry q of the serial order: (j, tuple index, own row?)
        auto entry = [&](int q, int &j, unsigned &ti) {
            const bool own = q >= rsplit && q < rsplit + len;
            const int kr = q < rsplit ? q : q - len; // position in the transposed row
            j = own ? s : __ldg(&so.rnbr[rbase + kr]);
            ti = own ? __ldg(&so.ctix[base + q - rsplit]) : __ldg(&so.crtix[rbase + kr]);
        };
        int jn = s; unsigned tin = 0;
        Rec tn{};
        double Pn[DIMS * DIMS];
        auto gather = [&]() {
            tn = ld_rec_ro(so.ttab + tin);
            const double *pp = so.PkA + 9 * (size_t)jn;
#pragma unroll
            for (int a = 0; a < DIMS; ++a)
#pragma unroll
                for (int b2 = 0; b2 < DIMS; ++b2) Pn[a * DIMS + b2] = pp[3 * a + b2];
        };
#pragma unroll
        for (int e = 0; e < DIMS * DIMS; ++e) Pn[e] = 0.0;
        if (lane < total) { entry(lane, jn, tin); gather(); }
        if (kTeam + lane < total) entry(kTeam + lane, jn, tin);
        int buf = 0;
        for (int c0 = 0; c0 < total; c0 += kTeam, buf ^= 1) {
            const Rec t = tn;
            double P[DIMS * DIMS];
#pragma unroll
            for (int e = 0; e < DIMS * DIMS; ++e) P[e] = Pn[e];
            if (c0 + kTeam + lane < total) gather();
            if (c0 + 2 * kTeam + lane < total) entry(c0 + 2 * kTeam + lane, jn, tin);
            const int q = c0 + lane;
            if (q < total) {
                const bool own = q >= rsplit && q < rsplit + len;
                const double d0[3] = {t.a, t.b, DIMS == 3 ? t.c : 0.0};
                for (int a = 0; a < DIMS; ++a) {
                    double f = 0.0;
                    for (int b2 = 0; b2 < DIMS; ++b2) f = add(f, mul(P[a * DIMS + b2], d0[b2]));
                    f = mul(f, t.d);
                    const double term = mul(mul(ir, f), edt);
                    Ts[team][buf][lane][a] = own ? term : -term; // :2883 v_s += ..., :2885 v_s -= ... (a - b == a + (-b) exactly)
                }
            }
            __syncwarp(mask);
            if (lane < DIMS) {
                const int m = total - c0;
                double tv[kTeam];
#pragma unroll
                for (int p = 0; p < kTeam; ++p) tv[p] = Ts[team][buf][p][lane];
#pragma unroll
                for (int p = 0; p < kTeam; ++p) v = p < m ? add(v, tv[p]) : v;
            }
        }
        if (lane < 3) Vs[team][lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < TEAMS) {
        const int s2 = s_lo + blockIdx.x * TEAMS + threadIdx.x;
        if (s2 < s_hi) {
            double v[3] = {Vs[threadIdx.x][0], Vs[threadIdx.x][1], Vs[threadIdx.x][2]};
            solid_pass2_finish<RING>(so, s2, v, W0, W1, W2, edt, module, double_update, ring);
        }
    }
    if (RING) ring_complete(ctl, 1, ring);
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
    int split = 0;
    for (int c = so.roff[s]; c < so.roff[s + 1]; ++c, ++kk) {
        const int j = so.rnbr[c]; // row j lists s: x0_js = Mod(x0_s - x0_j ...) as row j computes it
        const size_t k = (size_t)kk * ns + s;
        const double a = minimg_exact(xi0, so.x0[j], W0), b = minimg_exact(yi0, so.y0[j], W1);
        const double cc = DIMS == 3 ? minimg_exact(zi0, so.z0[j], W2) : 0.0;
        so.ernbr[k] = j;
        so.rd0x[k] = a; so.rd0y[k] = b; so.rd0z[k] = cc;
        so.rw[k] = tl_weight<DIMS>(a, b, cc, radius, cw);
        if (j < s) split = kk + 1; // (rows ascending: the entries of rows j < s come first)
    }
    so.rlen[s] = kk;
    so.rsplit[s] = split;
}

// ------------------------------------------------------------------------------------------------
// upload / download helpers (AoS original order <-> sorted SoA)
// host arrays (global, original order AoS) -> the slots this context holds; ids == nullptr: all, in order
__global__ void k_upload_split(int n, const int *__restrict__ ids, const int *__restrict__ type, const double *__restrict__ x3,
                               const double *__restrict__ v3, Particles p, int *__restrict__ solid_slot, int solid_base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids ? ids[i] : i;
    const size_t o = 3 * (size_t)id;
    p.x[i] = x3[o]; p.y[i] = x3[o + 1]; p.z[i] = x3[o + 2];
    p.vx[i] = v3[o]; p.vy[i] = v3[o + 1]; p.vz[i] = v3[o + 2];
    p.type[i] = type[id]; p.id[i] = id; p.key[i] = 0;
    if (solid_slot && is_structure_type(type[id])) solid_slot[id - solid_base] = i;
}
// device-side generator: the lattice fill of generator/generator.cpp:654-680.  The host computes the
// three axis tables of every cuboid exactly as the generator + its `%e` text do (a few hundred doubles); the particles --
// x outer, y, z inner, cuboid after cuboid -- are written straight into device memory.
struct GenCuboid {
    long long first;  // index of the cuboid's first particle
    int nx, ny, nz;   // lattice points per axis
    int ax, ay, az;   // offsets of its axis tables in `axes`
    int type;
    double v[3];
};
constexpr int kMaxCuboids = 64;
struct GenPlan { int count; GenCuboid c[kMaxCuboids]; };
__global__ void k_generate(long long n, GenPlan plan, const double *__restrict__ axes, int *__restrict__ type, double *__restrict__ x3,
                           double *__restrict__ x03, double *__restrict__ v3)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q = 0;
    while (q + 1 < plan.count && plan.c[q + 1].first <= i) ++q;
    const GenCuboid &g = plan.c[q];
    const long long r = i - g.first;
    const int iz = (int)(r % g.nz), iy = (int)((r / g.nz) % g.ny), ix = (int)(r / ((long long)g.nz * g.ny));
    const double x = axes[g.ax + ix], y = axes[g.ay + iy], z = axes[g.az + iz];
    const size_t o = 3 * (size_t)i;
    type[i] = g.type;
    x3[o] = x; x3[o + 1] = y; x3[o + 2] = z;
    x03[o] = x; x03[o + 1] = y; x03[o + 2] = z;
    v3[o] = g.v[0]; v3[o + 1] = g.v[1]; v3[o + 2] = g.v[2];
}

// device-side generator on a slab: which of the generated particles this slab keeps (all solids + the fluid / wall particles
// of its own columns: the rule of mphx_upload), and their ids once the mask is scanned
__global__ void k_generated_keep(long long n, const int *__restrict__ type, const double *__restrict__ x3, GridDesc g, int *__restrict__ keep)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool k = is_structure_type(type[i]);
    if (!k) {
        int cx = cell_coord_exact(x3[3 * (size_t)i], g.mn0g, g.cellw, g.nxg) - g.xoff; // :1671
        if (cx < 0) cx += g.nxg; else if (cx >= g.nxg) cx -= g.nxg;
        k = cx >= g.range && cx < g.nx - g.range;
    }
    keep[i] = k ? 1 : 0;
}
__global__ void k_generated_ids(long long n, const int *__restrict__ keep, const int *__restrict__ scan, int *__restrict__ ids)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep[i]) ids[scan[i]] = (int)i;
}

// the solids in their reference configuration as a particle set (for the initial-list build)
__global__ void k_solid_reference_particles(Solid so, Particles p)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    p.x[s] = so.x0[s]; p.y[s] = so.y0[s]; p.z[s] = so.z0[s];
    p.vx[s] = 0.0; p.vy[s] = 0.0; p.vz[s] = 0.0;
    p.type[s] = so.type[s]; p.id[s] = so.sb + s; p.key[s] = 0;
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
// ownership mask of the sorted slots (slab mode: ghosts, parked solids and solids owned by another
// slab's columns are not reported by this slab)
__global__ void k_owned_mask(int n, Particles p, GridDesc g, int solids_too, int *__restrict__ mask)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int k = p.key[q], t = p.type[q];
    bool own = !(t & kGhost);
    if (g.slab) {
        // solids are replicated: solids_too 0 = never, 1 = always (one reporting slab), 2 = the slab
        // that owns the particle's current column (bucket-related fields)
        if (is_structure_type(t) && solids_too != 2) own = own && solids_too;
        else own = own && k < g.ncells && column_owned(g, key_column(g, k));
    }
    mask[q] = own ? 1 : 0;
}
// owned fluid / wall particles per GLOBAL bucket column (the input of the slab re-balancing)
__global__ void k_column_histogram(const Ctl *ctl, Particles p, GridDesc g, int *__restrict__ hist)
{
    const int n = ctl->n;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int k = p.key[q], t = p.type[q];
        if ((t & kGhost) || is_structure_type(t) || k >= g.ncells) continue;
        const int col = key_column(g, k);
        if (!column_owned(g, col)) continue;
        int cx = col + g.xoff;
        if (cx < 0) cx += g.nxg; else if (cx >= g.nxg) cx -= g.nxg;
        atomicAdd(&hist[cx], 1);
    }
}
// compact state of the masked slots: out row r = scan[q] for mask[q] != 0; solids report the solid arrays
__global__ void k_compact_owned(int n, Particles p, Solid sol, const int *__restrict__ mask, const int *__restrict__ scan,
                                int *__restrict__ slot_of_row, int *__restrict__ ids, double *__restrict__ x3, double *__restrict__ v3)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || !mask[q]) return;
    const int r = scan[q];
    const int id = p.id[q];
    double x = p.x[q], y = p.y[q], z = p.z[q], vx = p.vx[q], vy = p.vy[q], vz = p.vz[q];
    if (is_structure_type(p.type[q])) {
        const int s = id - sol.sb;
        x = sol.x[s]; y = sol.y[s]; z = sol.z[s]; vx = sol.vx[s]; vy = sol.vy[s]; vz = sol.vz[s];
    }
    slot_of_row[r] = q;
    ids[r] = id;
    const size_t o = 3 * (size_t)r;
    x3[o] = x; x3[o + 1] = y; x3[o + 2] = z; v3[o] = vx; v3[o + 1] = vy; v3[o + 2] = vz;
}
// the inverse: rows (in the order of the last k_compact_owned) back into their slots; err |= 1 on an id mismatch
__global__ void k_scatter_owned(int count, Particles p, Solid sol, const int *__restrict__ slot_of_row, const int *__restrict__ ids,
                                const double *__restrict__ x3, const double *__restrict__ v3, int *__restrict__ err)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const int q = slot_of_row[r];
    if (p.id[q] != ids[r]) { atomicOr(err, 1); return; }
    const size_t o = 3 * (size_t)r;
    const double x = x3[o], y = x3[o + 1], z = x3[o + 2], vx = v3[o], vy = v3[o + 1], vz = v3[o + 2];
    p.x[q] = x; p.y[q] = y; p.z[q] = z; p.vx[q] = vx; p.vy[q] = vy; p.vz[q] = vz;
    if (is_structure_type(p.type[q])) {
        const int s = ids[r] - sol.sb;
        sol.x[s] = x; sol.y[s] = y; sol.z[s] = z; sol.vx[s] = vx; sol.vy[s] = vy; sol.vz[s] = vz;
    }
}
__global__ void k_global_keys(int n, Particles p, GridDesc g, int *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = p.key[q] < g.ncells ? global_key(g, p.key[q]) : -1;
}
__global__ void k_real_types(int n, Particles p, int *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = real_type(p.type[q]);
}
__global__ void k_gather_vec3(int n, const int *__restrict__ id, const int *__restrict__ mask, const double *__restrict__ a,
                              const double *__restrict__ b, const double *__restrict__ c, double *__restrict__ out3)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || !mask[q]) return;
    const size_t o = 3 * (size_t)id[q];
    out3[o] = a[q]; out3[o + 1] = b[q]; out3[o + 2] = c[q];
}
__global__ void k_gather_scalar(int n, const int *__restrict__ id, const int *__restrict__ mask, const double *__restrict__ a,
                                double *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || !mask[q]) return;
    out[id[q]] = a[q];
}
__global__ void k_gather_int(int n, const int *__restrict__ id, const int *__restrict__ mask, const int *__restrict__ a,
                             int *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || !mask[q]) return;
    out[id[q]] = a[q];
}
// solid arrays -> original-order AoS (overrides the stale sorted copies)
// (owner_type != nullptr: only the solids this slab evaluates -- kSolidOwned in the type of their slot)
__global__ void k_solid_vec3_to_orig(Solid so, const double *__restrict__ a, const double *__restrict__ b,
                                     const double *__restrict__ c, double *__restrict__ out3, const int *__restrict__ owner_type)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    if (owner_type && !(owner_type[so.slot[s]] & kSolidOwned)) return;
    const size_t o = 3 * (size_t)(so.sb + s);
    out3[o] = a[s]; out3[o + 1] = b[s]; out3[o + 2] = c[s];
}
__global__ void k_solid_scalar_to_orig(Solid so, const double *__restrict__ a, double *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= so.ns) return;
    out[so.sb + s] = a[s];
}
__global__ void k_solid_tensor_to_orig(Solid so, const double *__restrict__ M, double *__restrict__ out9, int s_lo, int s_hi)
{
    const int s = s_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s_hi) return;
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
