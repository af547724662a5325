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
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
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

static int parse_module(const char *s)
{
    if (!s) return MPHX_MODULE_BAR;
    if (!strcmp(s, "bar") || !strcmp(s, "Bar_Module") || !strcmp(s, "1")) return MPHX_MODULE_BAR;
    if (!strcmp(s, "dam") || !strcmp(s, "DAM_Module") || !strcmp(s, "2")) return MPHX_MODULE_DAM;
    if (!strcmp(s, "none") || !strcmp(s, "0")) return MPHX_MODULE_NONE;
    fprintf(stderr, "unknown module '%s' (bar|dam|none)\n", s);
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
    int rc = mphx_read_data_file(datafilename.c_str(), &p, &rc_, invalid_line, nullptr);
    if (rc) die("readDataFile", rc);
    int n = 0, *property = nullptr;
    double *position = nullptr, *initial_position = nullptr, *velocity = nullptr;
    rc = mphx_read_grid_file(gridfilename.c_str(), &p, &n, &property, &position, &initial_position, &velocity);
    if (rc) die("readGridFile", rc);
    {
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
    mphx_ctx *ctx = nullptr;
    rc = mphx_create(&ctx, &p, getenv("MPHX_DEVICE") ? atoi(getenv("MPHX_DEVICE")) : 0);
    if (rc) die("mphx_create", rc);
    mphx_constants k;
    mphx_get_constants(ctx, &k);
    log_printf("N0a = %e, count=%d\n", k.n0a, k.n0a_count); // :1258
    log_printf("N0p = %e, count=%d\n", k.n0p, k.n0p_count); // :1303
    if ((rc = mphx_upload(ctx, n, property, position, initial_position, velocity))) die("mphx_upload", rc);
    if ((rc = mphx_init(ctx))) die("mphx_init", rc);
    mphx_set_timing(ctx, 1);

    const size_t N = (size_t)n;
    std::vector<double> force(3 * N), accel(3 * N), stress(9 * N), strain(9 * N);
    std::vector<int> nbc(N), inbc(N);
    auto download_state = [&]() {
        mphx_host_views v;
        memset(&v, 0, sizeof(v));
        v.position = position;
        v.velocity = velocity;
        int e = mphx_download(ctx, &v);
        if (e) die("mphx_download", e);
    };
    auto write_vtk = [&](const char *fn) {
        mphx_host_views v;
        memset(&v, 0, sizeof(v));
        v.property = property; v.position = position; v.velocity = velocity;
        v.force = force.data(); v.acceleration = accel.data(); v.stress = stress.data(); v.strain = strain.data();
        v.neighbor_count = nbc.data(); v.initial_structure_neighbor_count = inbc.data();
        int e = mphx_download(ctx, &v);
        if (e) die("mphx_download", e);
        if ((e = mphx_write_vtk_file(fn, n, initial_position, &v))) die("writeVtkFile", e);
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
            download_state();
            if ((rc = mphx_write_prof_file(filename, Time, &p, n, property, position, initial_position, velocity)))
                die("writeProfFile", rc);
            log_printf("@ Prof Output Time : %e\n", Time);
            OutputNext += rc_.output_interval;
        }
        double t1 = now(); sOther += t1 - tFrom; tFrom = t1;
        if ((rc = mphx_step(ctx, 1))) die("mphx_step", rc); // :596-663
        if (Time + 1.0e-5 * Dt >= VtkOutputNext) {          // :672-683
            if ((rc = mphx_sync(ctx))) die("mphx_sync", rc);
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
    if ((rc = mphx_sync(ctx))) die("mphx_sync", rc);
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
        log_printf("virial calculation:      %lf [CPU sec]\n", 0.0);
        log_printf("other calculation:       %lf [CPU sec]\n", sOther);
        log_printf("total:                   %lf [CPU sec]\n", neigh + expl + sOther);
        log_printf("total (check):           %lf [CPU sec]\n", now() - tStart);
    }
    mphx_destroy(ctx);
    mphx_free_host(property); mphx_free_host(position); mphx_free_host(initial_position); mphx_free_host(velocity);
    if (g_log) fclose(g_log);
    return 0;
}
