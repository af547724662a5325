/* mphx.h -- C-ABI boundary of the B200-native explicit MPH / total-Lagrangian FSI step.
 *
 * The reference (Ryo1011gd/ParticleMethod_FSI) has no plugin/FFI interface: its hot path is 22
 * `static void f(void)` procedures on file-scope globals, called in fixed order by main()
 * (src/main.cpp:596-663).  The stable external contract is the CLI (src/main.cpp:501-508), the
 * file formats (.data :729-786, .grid/.prof :788-982, .vtk :984-1189) and that call order.  This
 * header is the extern-"C" layer a maintainer would bind instead of the OpenACC regions: plain
 * pointers and sizes, no C++/torch types, error codes instead of exit(1).
 *
 * Every entry point cites the reference interface it replaces.  Arrays crossing the boundary use
 * the reference's own layout: AoS `double[N][3]` vectors, `double[N][3][3]` tensors, `int[N]`
 * scalars, ORIGINAL (file) particle order.  The library owns all device memory (cell-sorted SoA).
 *
 * There is NO CPU fallback: without an sm_100 device every compute entry point returns
 * MPHX_ERR_NO_DEVICE.  The file-format and host-constant functions need no GPU.
 */
#ifndef MPHX_H_INCLUDED
#define MPHX_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

#define MPHX_TYPE_COUNT 6 /* src/main.cpp:68  (fluid 0-1, structure 2-3, wall 4-5; :69-74) */
#define MPHX_VERSION 100

/* ---- error codes ---------------------------------------------------------------------------- */
enum {
    MPHX_OK = 0,
    MPHX_ERR_INVALID = -1,    /* bad argument / state                                         */
    MPHX_ERR_NO_DEVICE = -2,  /* no CUDA device of compute capability 10.x (no CPU fallback)  */
    MPHX_ERR_CUDA = -3,       /* a CUDA call failed; see mphx_last_error()                    */
    MPHX_ERR_IO = -4,         /* file could not be opened / parsed  (src/errorfunc.cpp:19-31) */
    MPHX_ERR_NOMEM = -5,      /* host or device allocation failed   (src/errorfunc.cpp:8-17)  */
    MPHX_ERR_UNSUPPORTED = -6,/* configuration outside the supported envelope                 */
    MPHX_ERR_OVERFLOW = -7    /* neighbour capacity exceeded (reference only counts, :1766)   */
};

/* ---- parameters ----------------------------------------------------------------------------- */
/* clamp module: the reference selects it with a compile-time #define (src/main.cpp:54-59); here it is a
 * run-time field.  All six clamp variants of updateElasticPosition (:1910-2082) are covered. */
enum { MPHX_MODULE_NONE = 0,
       MPHX_MODULE_BAR = 1          /* :54  Bar_Module:   clamp x0[0] < 0.001                 :1919 */,
       MPHX_MODULE_DAM = 2          /* :55  DAM_Module:   clamp x0[1] < 0.002                 :1968 */,
       MPHX_MODULE_TUREK_HRON = 3   /* :56  Turek_Hron:   clamp x0[0] < 0.205, Force kept     :1944 */,
       MPHX_MODULE_ROLLING1 = 4     /* :57  Rolling1:     clamp x0[1] < 0.003                 :1992 */,
       MPHX_MODULE_HYDROELASTIC = 5 /* :59  Hydroelastic: clamp x0[0] < 0.01 or > 1.99        :2016 */,
       MPHX_MODULE_ROLLING2 = 6     /*      Rolling2:     clamp x0[1] > 0.3420, and the only variant WITHOUT the
                                            second position update of quirk Q1 (:2040-2069)         */ };
/* wall kinematics (mphx_params.wall_module): 0 = translate / rotate with the .data file's Wall6 / Wall7 rows while
 * Time < 0.2 (:3033-3070); 1 = the reference's `#define Rolling` (:2958-3031): the walls roll about z through their
 * centres, theta(t) = 2 deg * sin(2 pi t / 1.646 s), at every step */
enum { MPHX_WALL_DEFAULT = 0, MPHX_WALL_ROLLING = 1 };

/* ref_compat bits: reproduce reference quirks (SURVEY.md Q-list).  Default: all on. */
enum { MPHX_COMPAT_DOUBLE_UPDATE = 1 /* Q1: updateElasticPosition advances twice, :2070-2079 */ };

typedef struct mphx_params {
    int dim;            /* 2 or 3: `#define TWO_DIMENSIONAL` src/main.cpp:50                   */
    int clamp_module;   /* MPHX_MODULE_*                                                       */
    int ref_compat;     /* MPHX_COMPAT_* bits                                                  */
    int wall_module;    /* MPHX_WALL_*                                                         */
    double time0;             /* Time from line 1 of the grid file          :797               */
    double dt;                /* Dt                                          :743              */
    double elastic_dt;        /* ElasticDt                                   :744              */
    double particle_spacing;  /* l0, grid header                             :799-801          */
    double domain_min[3];     /*                                             :802-804          */
    double domain_max[3];
    double radius_ratio_a;    /* RadiusRatioA (= RadiusRatioG, :1193)         :748             */
    double radius_ratio_p;    /*                                             :750              */
    double radius_ratio_v;    /*                                             :751              */
    double density[MPHX_TYPE_COUNT];          /* :752 */
    double bulk_modulus[MPHX_TYPE_COUNT];     /* :753 */
    double bulk_viscosity[MPHX_TYPE_COUNT];   /* :754 */
    double shear_viscosity[MPHX_TYPE_COUNT];  /* :755 */
    double surface_tension[MPHX_TYPE_COUNT];  /* :756 (file gives types 0,1,4,5) */
    double young_modulus[MPHX_TYPE_COUNT];    /* :757 (file gives types 2..5)    */
    double poisson_ratio[MPHX_TYPE_COUNT];    /* :758 */
    double interaction_ratio[MPHX_TYPE_COUNT][MPHX_TYPE_COUNT]; /* :759-764 */
    double gravity[3];                        /* :765 */
    double wall_center[MPHX_TYPE_COUNT][3];   /* :766-767 (Wall6 -> type 4, Wall7 -> type 5) */
    double wall_velocity[MPHX_TYPE_COUNT][3];
    double wall_omega[MPHX_TYPE_COUNT][3];
} mphx_params;

/* what the driver (not the step) needs from the .data file: src/main.cpp:745-747 */
typedef struct mphx_run_control {
    double output_interval;     /* OutputInterval    */
    double vtk_output_interval; /* VtkOutputInterval */
    double end_time;            /* EndTime           */
} mphx_run_control;

/* host-side constants computed once by the reference's initialize* procedures; they must be
 * bit-identical to the reference (initializeWeight :1191-1309, initializeFluid :1312-1341,
 * initializeWall :1371-1410, initializeDomain :1412-1469). */
typedef struct mphx_constants {
    double particle_volume;                 /* :806-808 */
    double radius_a, radius_g, radius_p, radius_v, max_radius;
    double swa, swg, swp, swv, r2g, n0a, n0p;
    double cof_k, cof_a[MPHX_TYPE_COUNT];
    double wall_rotation[MPHX_TYPE_COUNT][3][3];
    double domain_max[3], domain_width[3];  /* after the integer-cell fix-up :1431-1440 */
    double cell_width;
    int cell_count[3];
    int cell_counts;                        /* product (int, like :1429)     */
    int n0a_count, n0p_count;               /* lattice points inside the radius (logged :1258,1303) */
    int stencil_range;                      /* ceil((MaxRadius+MARGIN)/CellWidth) :1744 */
} mphx_constants;

/* Host views for download: any pointer may be NULL (= not wanted).  Shapes as in the reference
 * (src/main.cpp:102-115,154-164,184-196); everything in ORIGINAL particle order. */
typedef struct mphx_host_views {
    int *property;        /* [N]       */
    double *position;     /* [N][3]    */
    double *velocity;     /* [N][3]    */
    double *force;        /* [N][3]    */
    double *acceleration; /* [N][3]    */
    double *pressure_p;   /* [N]       */
    double *vol_strain_p; /* [N]       */
    double *divergence_p; /* [N]       */
    double *density_a;    /* [N]       */
    double *gravity_center; /* [N][3]  */
    double *pressure_a;   /* [N]       */
    int *neighbor_count;  /* [N]  NeighborCount with the reference's bit-exact predicate :1764-1772 */
    int *initial_structure_neighbor_count; /* [N] :1608-1616 */
    int *cell_index;      /* [N]  CellId of the particle's bucket :1671-1674 */
    double *normalizer;       /* [N][3][3] */
    double *deform_gradient;  /* [N][3][3] */
    double *strain;           /* [N][3][3] */
    double *stress;           /* [N][3][3] */
    double *lambda_lames;     /* [N] */
    double *mu_lames;         /* [N] */
    /* calculateVirialStressAtParticle (src/main.cpp:3077-3318), evaluated on request over the state of the
     * last step (the reference calls it on VTK steps, :671-673) */
    double *virial_stress;    /* [N][3][3] VirialStressAtParticle   :190 */
    double *virial_pressure;  /* [N]       VirialPressureAtParticle :189 */
} mphx_host_views;

typedef struct mphx_ctx mphx_ctx;

/* ---- misc ----------------------------------------------------------------------------------- */
int mphx_version(void);
const char *mphx_strerror(int code);
/* last CUDA / IO diagnostic text of the calling thread ("" if none) */
const char *mphx_last_error(void);
/* number of visible sm_100 devices (0 when there is none or no driver) */
int mphx_device_count(void);
/* sizeof of the boundary structs, for FFI layout checks: 0 params, 1 run_control, 2 constants,
 * 3 host_views */
int mphx_abi_sizeof(int which);

/* ---- file formats (host only; replaces readDataFile/readGridFile/writeProfFile/writeVtkFile) -- */
void mphx_params_default(mphx_params *p, mphx_run_control *rc);
/* src/main.cpp:729-786.  Unknown lines are reported through `on_invalid_line` (may be NULL),
 * exactly the lines the reference logs as `Invalid line in data file`. */
int mphx_read_data_file(const char *filename, mphx_params *p, mphx_run_control *rc,
                        void (*on_invalid_line)(const char *line, void *user), void *user);
/* src/main.cpp:788-929.  Allocates the four arrays with malloc (free with mphx_free_host);
 * fills time0, particle_spacing, domain_* of *p. */
int mphx_read_grid_file(const char *filename, mphx_params *p, int *n_out, int **property,
                        double **position, double **initial_position, double **velocity);
void mphx_free_host(void *ptr);
/* src/main.cpp:957-982 (byte-identical text) */
int mphx_write_prof_file(const char *filename, double time, const mphx_params *p, int n,
                         const int *property, const double *position,
                         const double *initial_position, const double *velocity);
/* src/main.cpp:984-1189 (byte-identical text incl. the duplicated `velocity` section) */
int mphx_write_vtk_file(const char *filename, int n, const double *initial_position,
                        const mphx_host_views *fields);
/* lossless binary checkpoint: every double bit for bit plus the wall centres, which the 7-digit
 * .prof text (:973-978) cannot carry.  mphx_write_checkpoint takes host arrays (p->wall_center = the CURRENT centres,
 * see mphx_get_wall_centers); mphx_read_checkpoint allocates the four arrays with malloc (free with mphx_free_host) and
 * fills time0, dim, particle_spacing, domain_*, wall_center of *p, so that create/upload/init continue the run. */
int mphx_write_checkpoint(const char *filename, double time, const mphx_params *p, int n, const int *property,
                          const double *position, const double *initial_position, const double *velocity);
int mphx_read_checkpoint(const char *filename, mphx_params *p, int *n_out, int **property, double **position,
                         double **initial_position, double **velocity);
/* first..last index of each particle class, src/main.cpp:909-929; ranges[6] =
 * {FluidBegin,FluidEnd,StructureBegin,StructureEnd,WallBegin,WallEnd} (-1 when absent) */
void mphx_class_ranges(int n, const int *property, int ranges[6]);

/* ---- host constants (no GPU needed) ----------------------------------------------------------- */
/* initializeWeight/Fluid/Wall/Domain, src/main.cpp:1191-1469 */
int mphx_compute_constants(const mphx_params *p, mphx_constants *c);

/* ---- context life cycle ------------------------------------------------------------------------ */
/* replaces `acc enter data` (src/main.cpp:775-784, 830-891) */
int mphx_create(mphx_ctx **ctx, const mphx_params *p, int device);
/* replaces `acc exit data` (src/main.cpp:704-723) */
void mphx_destroy(mphx_ctx *ctx);
/* replaces `acc update device` (src/main.cpp:549-560, 946-950).  The caller keeps its arrays. */
int mphx_upload(mphx_ctx *ctx, int n, const int *property, const double *position,
                const double *initial_position, const double *velocity);
/* Device-side generator: the lattice fill of the reference's pre-processor
 * (generator/generator.cpp:654-680: per Cuboid, start at lower + spacing/2, accumulate while p < upper - 0.49 spacing,
 * x outer / y / z inner) written straight into device memory -- the particle set equals what the solver would read from
 * the generator's .grid text (`%e`, 7 digits), without the text and without host arrays (a 10^8-particle .grid is 12 GB).
 * Cuboids in file order; each particle class must be contiguous.  Replaces mphx_upload on a single context. */
typedef struct mphx_cuboid {
    int type;
    int reserved;
    double lower[3], upper[3], spacing, velocity[3];
} mphx_cuboid;
long long mphx_generate_count(const mphx_cuboid *cuboids, int ncuboids); /* particles the cuboids hold (-1: invalid) */
/* the pre-processor's own input file (generator/generator.cpp:127-262; StartCuboid blocks only): fills time0 (0),
 * particle_spacing and the domain through the `%e` text the generator would write, mallocs the cuboids (mphx_free_host) */
int mphx_read_boid_file(const char *filename, mphx_params *p, mphx_cuboid **cuboids, int *ncuboids);
int mphx_upload_generated(mphx_ctx *ctx, const mphx_cuboid *cuboids, int ncuboids); /* also on a configured slab: it keeps its share */
/* fluid / wall particles of the cuboids per bucket column, from the axis tables alone: cuts the slabs of a generated case */
int mphx_generate_column_histogram(const mphx_cuboid *cuboids, int ncuboids, double domain_min0, double cell_width, int ncols,
                                   long long *hist /* ncols */);
/* per-step variant of the above for a caller that keeps the state on the host: replaces Position
 * and Velocity only (original order, [N][3]); asynchronous on the context's stream, so pass
 * page-locked buffers and keep them alive until the next mphx_sync/mphx_download */
int mphx_upload_state(mphx_ctx *ctx, const double *position, const double *velocity);
/* replaces src/main.cpp:534-537 + 564-570 (initialize*, calculateInitialNeighbor, first
 * calculateNeighbor/DensityA/GravityCenter/DensityP, calculateLamesconstant, calculateNormalizer) */
int mphx_init(mphx_ctx *ctx);
int mphx_get_constants(const mphx_ctx *ctx, mphx_constants *c);
/* the wall centres as advanced so far (src/main.cpp:3066-3070: WallCenter += WallVelocity * Dt every step) */
int mphx_get_wall_centers(const mphx_ctx *ctx, double centers[MPHX_TYPE_COUNT][3]);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* `nsteps` iterations of the loop body src/main.cpp:596-663 followed by Time += Dt (:685).
 * Enqueues on the context's stream and returns; use mphx_sync (or mphx_download) to wait. */
int mphx_step(mphx_ctx *ctx, int nsteps);
/* debugging / stage parity: one step that stops after calculateConvection (:647), i.e. without
 * the solid sub-steps and without advancing Time */
int mphx_step_fluid_only(mphx_ctx *ctx);
int mphx_sync(mphx_ctx *ctx);
double mphx_time(const mphx_ctx *ctx);
int mphx_set_time(mphx_ctx *ctx, double time);
/* replaces `acc update host` (src/main.cpp:987-989); un-permutes to original particle order */
int mphx_download(mphx_ctx *ctx, const mphx_host_views *views);

/* compact per-context I/O (no reference counterpart: the reference keeps one host copy of everything).
 * A single context owns every particle; a slab context owns the fluid/wall particles of its columns and
 * reports ALL of the replicated solids.  Rows come in the context's current slot order.
 * mphx_upload_owned takes the rows of the LAST mphx_download_owned back (same ids in the same order,
 * values possibly modified; no step in between).  In slab mode every rank must be given the same solid
 * rows.  Buffers: ids[capacity], position/velocity[capacity][3]; page-locked memory makes the copies
 * asynchronous to the host. */
int mphx_download_owned(mphx_ctx *ctx, int capacity, int *ids, double *position, double *velocity, int *count);
int mphx_upload_owned(mphx_ctx *ctx, int count, const int *ids, const double *position, const double *velocity);

/* ---- parity / debug ------------------------------------------------------------------------------ */
/* The neighbour SETS of calculateNeighbor (src/main.cpp:1730-1810) with its bit-exact predicate,
 * as CSR in original particle ids, each row sorted ascending.  offsets has N+1 entries.
 * ids may be NULL to query the total size (returned through offsets[N]). */
int mphx_debug_neighbors(mphx_ctx *ctx, long long *offsets, int *ids, long long ids_capacity);
/* InitialStructureNeighbor of calculateInitialNeighbor (src/main.cpp:1497-1644), rows sorted */
int mphx_debug_initial_structure_neighbors(mphx_ctx *ctx, long long *offsets, int *ids,
                                           long long ids_capacity);

/* `nsteps` steps bracketed by CUDA events on the context's stream (the stream the kernels are
 * launched on); returns the device time of the whole region in milliseconds after synchronising. */
int mphx_timed_steps(mphx_ctx *ctx, int nsteps, double *elapsed_ms);
/* per-phase device timers (CUDA events on the context's stream around each phase of every step):
 * off by default; switching on resets the accumulators */
int mphx_set_timing(mphx_ctx *ctx, int on);
/* accumulated device milliseconds: [0] bucket rebuild (the reference's "neighbor search" timer,
 * src/main.cpp:611), [1] pass 1, [2] pass 2, [3] solid sub-steps ([1]+[2]+[3] = its "explicit
 * calculation" timer, :669) */
int mphx_get_timers(mphx_ctx *ctx, double ms[4]);
/* the same split by kernel group: [0] bucket rebuild, [1] candidate filter (k_filter), [2] pass 1 over the
 * candidate list, [3] pass 2 over the candidate list (+ integration), [4] solid sub-steps */
int mphx_get_kernel_timers(mphx_ctx *ctx, double ms[5]);
/* Device-side timeline (no reference counterpart; the reference has wall-clock phase timers only, src/main.cpp:695-700):
   capacity > 0 switches on (code, %globaltimer ns) marks written by the wait / push / marker kernels of the following
   steps, 0 switches them off; mphx_trace_read returns the marks recorded so far (out[2i] = code, out[2i+1] = ns). */
int mphx_trace_enable(mphx_ctx *ctx, int capacity);
int mphx_trace_read(mphx_ctx *ctx, unsigned long long *out, int max_marks, int *count);
/* roofline helpers (measurement only): dense FP64 FMA throughput of a device in TFLOP/s (CUDA events around a pure DFMA
 * kernel), and the candidates / in-radius pairs of the current lists (out[0], out[1]) -- a sweep's algorithmic FP64 work
 * is ~15 flop per candidate examined + ~45 per in-radius pair */
int mphx_measure_fp64_peak(int device, double *tflops);
int mphx_count_pairs(mphx_ctx *ctx, unsigned long long out[2]);
/* accumulated device milliseconds of calculateVirialStressAtParticle (the reference's "virial calculation" timer, :674) */
double mphx_get_virial_ms(const mphx_ctx *ctx);
/* the solid sub-steps normally run on a second stream, overlapping pass 2 and the start of the next step;
 * on = 0 serialises them on the context's stream (isolated per-kernel timings), on = 1 restores the default */
int mphx_set_overlap(mphx_ctx *ctx, int on);
/* make the context's stream wait (device-side, no host synchronisation) for everything the library has
 * enqueued on its internal second stream: call before recording an end-of-region event on that stream */
int mphx_join(mphx_ctx *ctx);
/* number of kernel launches issued by mphx_step since mphx_create (for bench.py gpu_launches) */
long long mphx_launch_count(const mphx_ctx *ctx);
/* algorithmic HBM bytes of one step for the resident case (SURVEY.md 8(d) model):
 * N_f*368 + N_w*260 + N_s*(344+384*n_sub) */
double mphx_algorithmic_bytes_per_step(const mphx_ctx *ctx);

/* ---- candidate-list reuse / health -------------------------------------------------------------------------
 * The reference rebuilds its neighbour lists every step (its Verlet-skin variant, neighborCalculation :1472-1494, is
 * disabled at :607-609).  Here the CANDIDATE list (a conservative superset; every pair is still tested in fp64 with
 * the reference's cut-offs, so bucket ids, neighbour sets and all sums keep their contracts) is built with
 * radius + skin and reused until a particle has moved skin/2 -- decided on the device, no host round trip.
 * on = 0: rebuild every step.  skin in particle spacings (<= 0: keep the default 0.25); before mphx_upload only. */
int mphx_set_list_reuse(mphx_ctx *ctx, int on, double skin);
/* out[0] error flags (1 particle crossed more than a halo width, 2 message buffer too small, 4 bad arrival, 8 peer
 * timeout, 16 slots exhausted, 32 a non-finite position appeared), [1] slots held, [2] lists built, [3] steps that
 * reused a list, [4] steps the current list has served, [5] the current list carries the skin, [6] several solids
 * share a bucket of the reference configuration (solid path then 1e-10 instead of bit for bit), [7] ghosts held.
 * Synchronises the context. */
int mphx_get_status(mphx_ctx *ctx, int out[8]);

/* ---- multi-GPU: one context per x-slab (the reference has no distributed path) ---
 * The exchange lives in the library and is device-side: every context owns a mailbox that its peers write
 * into over NVLink (peer access inside one process, CUDA IPC between processes); mphx_step on a connected slab
 * context enqueues pre-step, votes, migration / halo, buckets, pass 1, PressureP exchange, pass 2, sub-steps
 * without any host synchronisation.  Every rank must issue the same calls (mphx_init, mphx_step) concurrently.
 *   mphx_slab_configure -> mphx_upload -> mphx_slab_mailbox -> (carry the handles) -> mphx_slab_connect -> mphx_init */
int mphx_set_stream(mphx_ctx *ctx, void *cuda_stream);
int mphx_partition_columns(const long long *hist, int ncols, int nranks, int halo, int *cuts /* nranks + 1 */);
int mphx_slab_configure(mphx_ctx *ctx, int rank, int nranks, int col_lo, int col_hi, int capacity,
                        int msg_capacity);
/* ipc_handle: 64 bytes (cudaIpcMemHandle_t); any output may be NULL */
int mphx_slab_mailbox(mphx_ctx *ctx, void *ipc_handle, void **device_ptr, long long *bytes);
/* ipc_handles: nranks * 64 bytes or NULL; device_ptrs[r] != NULL: rank r's mailbox lives in this process;
 * devices[r]: CUDA ordinal of rank r's GPU in this process (or NULL) */
int mphx_slab_connect(mphx_ctx *ctx, const void *ipc_handles, void *const *device_ptrs, const int *devices);
int mphx_slab_info(mphx_ctx *ctx, int out[4]);
/* In-place re-balancing of a running ring: sum mphx_slab_column_histogram over the ranks, let
 * mphx_rebalance_cuts move every interior cut towards the balanced position by at most one halo width, and hand every
 * rank its new columns with mphx_slab_recut before the same step -- the particles that change owner travel with that
 * step's ordinary migration (no host gather).  mphx_slab_columns: the columns a slab owns (out[0], out[1]). */
int mphx_slab_column_histogram(mphx_ctx *ctx, long long *hist /* ncols */, int ncols);
int mphx_rebalance_cuts(const long long *hist, int ncols, int nranks, int halo, const int *old_cuts, int *new_cuts, int *moved);
int mphx_slab_recut(mphx_ctx *ctx, int col_lo, int col_hi);
int mphx_slab_columns(mphx_ctx *ctx, int out[2]);

/* one process, N devices -- what the reference's main() needs to drive a whole box (csrc/main.cpp, MPHX_NGPU):
 * same life cycle as a single context; devices may be NULL (0..ndev-1) and may repeat */
typedef struct mphx_multi mphx_multi;
int mphx_multi_create(mphx_multi **m, const mphx_params *p, int ndev, const int *devices);
void mphx_multi_destroy(mphx_multi *m);
int mphx_multi_count(const mphx_multi *m);
mphx_ctx *mphx_multi_context(mphx_multi *m, int i);
int mphx_multi_upload(mphx_multi *m, int n, const int *property, const double *position,
                      const double *initial_position, const double *velocity);
int mphx_multi_upload_generated(mphx_multi *m, const mphx_cuboid *cuboids, int ncuboids); /* no particle array on the host */
int mphx_multi_init(mphx_multi *m);
int mphx_multi_step(mphx_multi *m, int nsteps);
int mphx_multi_sync(mphx_multi *m);
double mphx_multi_time(const mphx_multi *m);
int mphx_multi_download(mphx_multi *m, const mphx_host_views *views);
int mphx_multi_timed_steps(mphx_multi *m, int nsteps, double *elapsed_ms);
int mphx_multi_rebalance(mphx_multi *m, int *moved /* interior cuts that moved; may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* MPHX_H_INCLUDED */
