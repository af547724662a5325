// main.cpp -- drop-in driver with the reference's command line (src/main.cpp:490-727):
//
//   Mph_Elastic_Explicit <data> <grid> <prof%03d> <vtk%03d> <log> <nthreads> [dim] [module]
//
// It reads the generator's .grid and the solver's .data files, runs the explicit step on one B200
// through the extern-"C" layer of include/mphx.h, and writes the reference's .prof / .vtk / .log
// outputs in the same order and under the same names (quirk Q8: .prof holds the state BEFORE the
// step, .vtk the state AFTER it, both named with the same step index; `output.vtk` once before the
// loop).  The reference fixes dimension and clamp module at compile time (src/main.cpp:50,54-55);
// here they are the optional 7th/8th arguments or MPHX_DIM / MPHX_MODULE (defaults: 2, bar -- the
// shipped configuration).  `nthreads` is accepted and ignored (there is no CPU path).
// A grid argument ending in .boid (the pre-processor's input) generates the particles on the device(s) instead of reading a .grid.
// MPHX_REBALANCE_EVERY=K re-cuts the slabs of a multi-GPU run every K steps (mphx_multi_rebalance).
// MPHX_NGPU=N (2..16) runs the case on N devices of the box as x-slabs (mphx_multi_*: one process, the
// slabs exchange over NVLink inside the library); MPHX_DEVICE picks the device of a single-GPU run.
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mphx.h"

static FILE *g_log = nullptr;

static int log_printf(const char *fmt, ...) // tee to log file and stderr, src/log.cpp:19-30
{
    va_list a1, a2;
    va_start(a1, fmt);
    int r = g_log ? vfprintf(g_log, fmt, a1) : 0;
    va_end(a1);
    va_start(a2, fmt);
    r = vfprintf(stderr, fmt, a2);
    va_end(a2);
    return r;
}

static void invalid_line(const char *line, void *) { log_printf("Invalid line in data file \"%s\"\n", line); }

static void die(const char *what, int rc)
{
    log_printf("mphx: %s failed: %s (%s)\n", what, mphx_strerror(rc), mphx_last_error());
    exit(1);
}

// the .prof / .vtk text is formatted and written by a background thread from a snapshot of
// the downloaded arrays while the GPU keeps stepping (at 10^7 particles the ASCII writers of the reference,
// src/main.cpp:957-1189, take far longer than a step).  One job may wait while one is being written; a
// third output request blocks the main loop (bounded memory).  Same writer functions, same bytes, same
// file order.  MPHX_SYNC_IO=1 writes in the main thread instead.
class Writer {
  public:
    explicit Writer(bool async) : async_(async) { if (async_) th_ = std::thread([this] { run(); }); }
    ~Writer() { finish(); }
    void submit(std::function<void()> job)
    {
        if (!async_) { job(); return; }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return q_.size() < 1; });
        q_.push_back(std::move(job));
        cv_.notify_all();
    }
    void finish()
    {
        if (!async_ || !th_.joinable()) return;
        { std::lock_guard<std::mutex> lk(m_); done_ = true; }
        cv_.notify_all();
        th_.join();
    }

  private:
    void run()
    {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return done_ || !q_.empty(); });
                if (q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
                cv_.notify_all();
            }
            job();
        }
    }
    bool async_, done_ = false;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
};

// a snapshot of everything one output file needs (the writer thread owns it)
struct Snapshot {
    std::vector<int> property, nbc, inbc;
    std::vector<double> position, velocity, force, accel, stress, strain, virial, virialp;
};

static int parse_module(const char *s)
{
    if (!s) return MPHX_MODULE_BAR;
    if (!strcmp(s, "bar") || !strcmp(s, "Bar_Module") || !strcmp(s, "1")) return MPHX_MODULE_BAR;
    if (!strcmp(s, "dam") || !strcmp(s, "DAM_Module") || !strcmp(s, "2")) return MPHX_MODULE_DAM;
    if (!strcmp(s, "turek") || !strcmp(s, "Turek_Hron") || !strcmp(s, "3")) return MPHX_MODULE_TUREK_HRON;
    if (!strcmp(s, "rolling1") || !strcmp(s, "Rolling1") || !strcmp(s, "4")) return MPHX_MODULE_ROLLING1;
    if (!strcmp(s, "hydroelastic") || !strcmp(s, "Hydroelastic") || !strcmp(s, "5")) return MPHX_MODULE_HYDROELASTIC;
    if (!strcmp(s, "rolling2") || !strcmp(s, "Rolling2") || !strcmp(s, "6")) return MPHX_MODULE_ROLLING2;
    if (!strcmp(s, "none") || !strcmp(s, "0")) return MPHX_MODULE_NONE;
    fprintf(stderr, "unknown module '%s' (bar|dam|turek|rolling1|hydroelastic|rolling2|none)\n", s);
    exit(1);
}

int main(int argc, char *argv[])
{
    std::string logfilename = "sample.log", datafilename = "sample.data", gridfilename = "sample.grid";
    std::string proffilename = "sample%03d.prof", vtkfilename = "sample%03d.vtk"; // :76-80
    if (argc > 1) datafilename = argv[1];
    if (argc > 2) gridfilename = argv[2];
    if (argc > 3) proffilename = argv[3];
    if (argc > 4) vtkfilename = argv[4];
    if (argc > 5) logfilename = argv[5];
    const char *dim_s = argc > 7 ? argv[7] : getenv("MPHX_DIM");
    const char *mod_s = argc > 8 ? argv[8] : getenv("MPHX_MODULE");

    g_log = fopen(logfilename.c_str(), "w");
    if (!g_log) fprintf(stderr, "error in open %s\n", logfilename.c_str());
    {
        time_t t = time(NULL);
        log_printf("start reading files at %s\n", ctime(&t));
    }
    mphx_params p;
    mphx_run_control rc_;
    mphx_params_default(&p, &rc_);
    if (dim_s) p.dim = atoi(dim_s);
    p.clamp_module = parse_module(mod_s);
    if (getenv("MPHX_WALL") && !strcmp(getenv("MPHX_WALL"), "rolling")) p.wall_module = MPHX_WALL_ROLLING; // `#define Rolling`
    int rc = mphx_read_data_file(datafilename.c_str(), &p, &rc_, invalid_line, nullptr);
    if (rc) die("readDataFile", rc);
    int n = 0, *property = nullptr;
    double *position = nullptr, *initial_position = nullptr, *velocity = nullptr;
    // a grid argument ending in .ckp is a lossless checkpoint of this driver (MPHX_CHECKPOINT), not the reference's text;
    // one ending in .boid is the PRE-PROCESSOR's input: the lattice is generated on the device(s) (mphx_upload_generated,
    // SURVEY 8(f) N4) -- the run equals the one from the generator's .grid without that (12 GB at 10^8 particles) text
    const bool restart = gridfilename.size() > 4 && gridfilename.compare(gridfilename.size() - 4, 4, ".ckp") == 0;
    const bool from_boid = gridfilename.size() > 5 && gridfilename.compare(gridfilename.size() - 5, 5, ".boid") == 0;
    mphx_cuboid *cuboids = nullptr;
    int ncuboids = 0;
    if (from_boid) {
        if ((rc = mphx_read_boid_file(gridfilename.c_str(), &p, &cuboids, &ncuboids))) die("readBoidFile", rc);
        const long long total = mphx_generate_count(cuboids, ncuboids);
        if (total <= 0 || total > 0x7fffffffLL) die("mphx_generate_count", MPHX_ERR_INVALID);
        n = (int)total;
    } else {
        rc = restart ? mphx_read_checkpoint(gridfilename.c_str(), &p, &n, &property, &position, &initial_position, &velocity)
                     : mphx_read_grid_file(gridfilename.c_str(), &p, &n, &property, &position, &initial_position, &velocity);
        if (rc) die("readGridFile", rc);
    }
    if (!from_boid) {
        int r[6];
        mphx_class_ranges(n, property, r); // :931-944
        printf("Fluid Particles: %d\n", r[0] != -1 ? r[1] - r[0] : 0);
        printf("Structure Particles: %d\n", r[2] != -1 ? r[3] - r[2] : 0);
        printf("Wall Particles: %d\n", r[4] != -1 ? r[5] - r[4] : 0);
    }
    {
        time_t t = time(NULL);
        log_printf("start initialization at %s\n", ctime(&t));
    }
    const int ngpu = getenv("MPHX_NGPU") ? atoi(getenv("MPHX_NGPU")) : 1;
    mphx_ctx *ctx = nullptr;    // single context, or slab 0 of the multi-GPU run (constants, timers)
    mphx_multi *multi = nullptr;
    if (ngpu > 1) {
        // MPHX_DEVICES=0,1,2,... picks the devices (default 0..N-1; a device may repeat: slabs sharing one GPU)
        std::vector<int> devs;
        if (const char *ds = getenv("MPHX_DEVICES"))
            for (const char *q = ds; *q;) { devs.push_back(atoi(q)); q = strchr(q, ','); if (!q) break; ++q; }
        if (!devs.empty() && (int)devs.size() != ngpu) { log_printf("MPHX_DEVICES must list MPHX_NGPU devices\n"); exit(1); }
        if ((rc = mphx_multi_create(&multi, &p, ngpu, devs.empty() ? nullptr : devs.data()))) die("mphx_multi_create", rc);
        ctx = mphx_multi_context(multi, 0);
    } else {
        rc = mphx_create(&ctx, &p, getenv("MPHX_DEVICE") ? atoi(getenv("MPHX_DEVICE")) : 0);
        if (rc) die("mphx_create", rc);
    }
    mphx_constants k;
    mphx_get_constants(ctx, &k);
    log_printf("N0a = %e, count=%d\n", k.n0a, k.n0a_count); // :1258
    log_printf("N0p = %e, count=%d\n", k.n0p, k.n0p_count); // :1303
    if (multi) {
        if (from_boid) { if ((rc = mphx_multi_upload_generated(multi, cuboids, ncuboids))) die("mphx_multi_upload_generated", rc); }
        else if ((rc = mphx_multi_upload(multi, n, property, position, initial_position, velocity))) die("mphx_multi_upload", rc);
        if ((rc = mphx_multi_init(multi))) die("mphx_multi_init", rc);
    } else {
        if (from_boid) { if ((rc = mphx_upload_generated(ctx, cuboids, ncuboids))) die("mphx_upload_generated", rc); }
        else if ((rc = mphx_upload(ctx, n, property, position, initial_position, velocity))) die("mphx_upload", rc);
        if ((rc = mphx_init(ctx))) die("mphx_init", rc);
    }
    if (from_boid) { // the host copies the writers need (Property, InitialPosition) come back from the device(s)
        const size_t nn = (size_t)n;
        property = (int *)calloc(nn, sizeof(int));
        position = (double *)calloc(3 * nn, sizeof(double));
        initial_position = (double *)calloc(3 * nn, sizeof(double));
        velocity = (double *)calloc(3 * nn, sizeof(double));
        if (!property || !position || !initial_position || !velocity) die("malloc", MPHX_ERR_NOMEM);
        mphx_host_views v0;
        memset(&v0, 0, sizeof(v0));
        v0.property = property; v0.position = position; v0.velocity = velocity;
        if ((rc = multi ? mphx_multi_download(multi, &v0) : mphx_download(ctx, &v0))) die("mphx_download", rc);
        memcpy(initial_position, position, sizeof(double) * 3 * nn);
        mphx_free_host(cuboids);
        int r[6];
        mphx_class_ranges(n, property, r); // :931-944
        printf("Fluid Particles: %d\n", r[0] != -1 ? r[1] - r[0] : 0);
        printf("Structure Particles: %d\n", r[2] != -1 ? r[3] - r[2] : 0);
        printf("Wall Particles: %d\n", r[4] != -1 ? r[5] - r[4] : 0);
    }
    mphx_set_timing(ctx, 1);
    // MPHX_REBALANCE_EVERY=K (multi-GPU): re-cut the slabs on the current particle distribution every K steps
    const int rebalance_every = (multi && getenv("MPHX_REBALANCE_EVERY")) ? atoi(getenv("MPHX_REBALANCE_EVERY")) : 0;
    long long steps_taken = 0;
    auto do_step = [&]() {
        if (!multi) return mphx_step(ctx, 1);
        if (rebalance_every > 0 && steps_taken > 0 && steps_taken % rebalance_every == 0) {
            int moved = 0;
            const int r = mphx_multi_rebalance(multi, &moved);
            if (r) return r;
            if (moved) log_printf("re-balanced the slabs: %d cuts moved\n", moved);
        }
        ++steps_taken;
        return mphx_multi_step(multi, 1);
    };
    auto do_sync = [&]() { return multi ? mphx_multi_sync(multi) : mphx_sync(ctx); };
    auto do_download = [&](const mphx_host_views *v) { return multi ? mphx_multi_download(multi, v) : mphx_download(ctx, v); };

    const size_t N = (size_t)n;
    Writer writer(!(getenv("MPHX_SYNC_IO") && atoi(getenv("MPHX_SYNC_IO")) != 0));
    int iStepNow = 0;
    auto write_prof = [&](const std::string &fn, double time) { // state BEFORE the step (Q8)
        auto snap = std::make_shared<Snapshot>();
        snap->position.resize(3 * N); snap->velocity.resize(3 * N);
        mphx_host_views v;
        memset(&v, 0, sizeof(v));
        v.position = snap->position.data();
        v.velocity = snap->velocity.data();
        int e = do_download(&v);
        if (e) die("mphx_download", e);
        mphx_params pc = p; // (with the wall centres as advanced so far: the checkpoint carries them)
        mphx_get_wall_centers(ctx, pc.wall_center);
        const char *ckp = getenv("MPHX_CHECKPOINT"); // e.g. run%03d.ckp: a lossless restart file next to every .prof
        const std::string ckpfn = ckp ? [&] { char b[2048]; snprintf(b, sizeof(b), ckp, iStepNow); return std::string(b); }() : std::string();
        writer.submit([=, &p]() {
            int e2 = mphx_write_prof_file(fn.c_str(), time, &p, n, property, snap->position.data(), initial_position, snap->velocity.data());
            if (e2) die("writeProfFile", e2);
            if (!ckpfn.empty()) {
                e2 = mphx_write_checkpoint(ckpfn.c_str(), time, &pc, n, property, snap->position.data(), initial_position, snap->velocity.data());
                if (e2) die("writeCheckpoint", e2);
            }
        });
    };
    // the reference computes the virial stress on VTK steps but keeps its VTK sections commented out; MPHX_VTK_VIRIAL=1
    // evaluates it (single GPU) and writes those sections
    const bool vtk_virial = getenv("MPHX_VTK_VIRIAL") && atoi(getenv("MPHX_VTK_VIRIAL")) != 0 && !multi;
    auto write_vtk = [&](const std::string &fn) {
        auto snap = std::make_shared<Snapshot>();
        snap->property.resize(N); snap->nbc.resize(N); snap->inbc.resize(N);
        snap->position.resize(3 * N); snap->velocity.resize(3 * N); snap->force.resize(3 * N); snap->accel.resize(3 * N);
        snap->stress.resize(9 * N); snap->strain.resize(9 * N);
        mphx_host_views v;
        memset(&v, 0, sizeof(v));
        v.property = snap->property.data(); v.position = snap->position.data(); v.velocity = snap->velocity.data();
        v.force = snap->force.data(); v.acceleration = snap->accel.data(); v.stress = snap->stress.data(); v.strain = snap->strain.data();
        v.neighbor_count = snap->nbc.data(); v.initial_structure_neighbor_count = snap->inbc.data();
        if (vtk_virial) { // calculateVirialStressAtParticle on output steps (:671-673) and its VTK sections (:1128-1143)
            snap->virial.resize(9 * N); snap->virialp.resize(N);
            v.virial_stress = snap->virial.data(); v.virial_pressure = snap->virialp.data();
        }
        int e = do_download(&v);
        if (e) die("mphx_download", e);
        writer.submit([=]() {
            mphx_host_views w = v; // (the pointers stay valid: the snapshot lives as long as this job)
            (void)snap;
            int e2 = mphx_write_vtk_file(fn.c_str(), n, initial_position, &w);
            if (e2) die("writeVtkFile", e2);
        });
    };
    write_vtk("output.vtk"); // :572
    {
        time_t t = time(NULL);
        log_printf("start main roop at %s\n", ctime(&t));
    }
    double Time = p.time0, OutputNext = 0.0, VtkOutputNext = 0.0; // :86-90
    const double Dt = p.dt;
    int iStep = (int)(Time / Dt); // :578
    double sOther = 0.0, sStep = 0.0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tStart = now();
    double tFrom = tStart;
    while (Time < rc_.end_time + 1.0e-5 * Dt) { // :581
        if (Time + 1.0e-5 * Dt >= OutputNext) { // :583-589
            char filename[2048];
            snprintf(filename, sizeof(filename), proffilename.c_str(), iStep);
            iStepNow = iStep;
            write_prof(filename, Time);
            log_printf("@ Prof Output Time : %e\n", Time);
            OutputNext += rc_.output_interval;
        }
        double t1 = now(); sOther += t1 - tFrom; tFrom = t1;
        if ((rc = do_step())) die("mphx_step", rc); // :596-663
        if (Time + 1.0e-5 * Dt >= VtkOutputNext) {          // :672-683
            if ((rc = do_sync())) die("mphx_sync", rc);
            t1 = now(); sStep += t1 - tFrom; tFrom = t1;
            char filename[2048];
            snprintf(filename, sizeof(filename), vtkfilename.c_str(), iStep);
            write_vtk(filename);
            log_printf("@ Vtk Output Time : %e\n", Time);
            VtkOutputNext += rc_.vtk_output_interval;
            t1 = now(); sOther += t1 - tFrom; tFrom = t1;
        }
        Time += Dt; // :685
        iStep++;
    }
    if ((rc = do_sync())) die("mphx_sync", rc);
    writer.finish(); // every output file is complete before the timers are logged
    {
        double t1 = now(); sStep += t1 - tFrom;
        double ms[4] = {0, 0, 0, 0};
        mphx_get_timers(ctx, ms);
        time_t t = time(NULL);
        log_printf("end main roop at %s\n", ctime(&t));
        // same six lines as src/main.cpp:695-700; times are wall/device seconds of this process
        const double neigh = ms[0] * 1e-3, expl = (ms[0] > 0 ? (ms[1] + ms[2] + ms[3]) * 1e-3 : sStep);
        log_printf("neighbor search:         %lf [CPU sec]\n", neigh);
        log_printf("explicit calculation:    %lf [CPU sec]\n", expl);
        log_printf("virial calculation:      %lf [CPU sec]\n", mphx_get_virial_ms(ctx) * 1e-3);
        log_printf("other calculation:       %lf [CPU sec]\n", sOther);
        log_printf("total:                   %lf [CPU sec]\n", neigh + expl + sOther);
        log_printf("total (check):           %lf [CPU sec]\n", now() - tStart);
    }
    if (multi) mphx_multi_destroy(multi); else mphx_destroy(ctx);
    mphx_free_host(property); mphx_free_host(position); mphx_free_host(initial_position); mphx_free_host(velocity);
    if (g_log) fclose(g_log);
    return 0;
}
