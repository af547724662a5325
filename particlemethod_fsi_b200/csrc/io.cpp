// io.cpp -- the reference's file formats at the drop-in boundary (host only, no GPU needed).
//
//   .data  key/value text     readDataFile   src/main.cpp:729-786
//   .grid  particle text      readGridFile   src/main.cpp:788-929   (written by generator.cpp:839-862)
//   .prof  same layout        writeProfFile  src/main.cpp:957-982
//   .vtk   legacy ASCII       writeVtkFile   src/main.cpp:984-1189
//
// Output text must be byte-identical to the reference's, so the printf conversions (%e, %d, the
// float casts of the VTK fields, the duplicated `velocity` block, the blank lines) are kept.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mphx.h"
#include "mphx_internal.h"

extern "C" void mphx_params_default(mphx_params *p, mphx_run_control *rc)
{
    if (p) {
        std::memset(p, 0, sizeof(*p));
        p->dim = 2;                      // the shipped build defines TWO_DIMENSIONAL (:50)
        p->clamp_module = MPHX_MODULE_BAR; // and Bar_Module (:54)
        p->ref_compat = MPHX_COMPAT_DOUBLE_UPDATE;
        p->dt = 1.0e100;         // :91
        p->elastic_dt = 1.0e100; // :92
    }
    if (rc) std::memset(rc, 0, sizeof(*rc));
}

namespace {

struct DataKey {
    const char *key;
    int count;
    double *dst[9];
};

// one `else if (sscanf(buf, " Key %lf ...") == count)` arm of :743-767
bool match_key(const char *line, const DataKey &k)
{
    std::string fmt = " ";
    fmt += k.key;
    for (int i = 0; i < k.count; ++i) fmt += " %lf";
    double *d = const_cast<double **>(k.dst)[0];
    (void)d;
    int got = 0;
    switch (k.count) {
    case 1: got = std::sscanf(line, fmt.c_str(), k.dst[0]); break;
    case 3: got = std::sscanf(line, fmt.c_str(), k.dst[0], k.dst[1], k.dst[2]); break;
    case 4: got = std::sscanf(line, fmt.c_str(), k.dst[0], k.dst[1], k.dst[2], k.dst[3]); break;
    case 6: got = std::sscanf(line, fmt.c_str(), k.dst[0], k.dst[1], k.dst[2], k.dst[3], k.dst[4], k.dst[5]); break;
    default: return false;
    }
    return got == k.count;
}

bool match_wall(const char *line, const char *name, mphx_params *p, int t)
{
    std::string fmt = std::string(" ") + name + "  Center %lf %lf %lf Velocity %lf %lf %lf Omega %lf %lf %lf";
    double *c = p->wall_center[t], *v = p->wall_velocity[t], *w = p->wall_omega[t];
    return std::sscanf(line, fmt.c_str(), &c[0], &c[1], &c[2], &v[0], &v[1], &v[2], &w[0], &w[1], &w[2]) == 9;
}

} // namespace

extern "C" int mphx_read_data_file(const char *filename, mphx_params *p, mphx_run_control *rc,
                                   void (*on_invalid_line)(const char *, void *), void *user)
{
    if (!filename || !p || !rc) return MPHX_ERR_INVALID;
    FILE *fp = std::fopen(filename, "r");
    if (!fp) {
        mphx::set_last_error(std::string("cannot open data file ") + filename);
        return MPHX_ERR_IO;
    }
    double *D = p->density, *K = p->bulk_modulus, *L = p->bulk_viscosity, *M = p->shear_viscosity;
    double *S = p->surface_tension, *Y = p->young_modulus, *Po = p->poisson_ratio;
    std::vector<DataKey> keys = {
        {"Dt", 1, {&p->dt}},
        {"ElasticDt", 1, {&p->elastic_dt}},
        {"OutputInterval", 1, {&rc->output_interval}},
        {"VtkOutputInterval", 1, {&rc->vtk_output_interval}},
        {"EndTime", 1, {&rc->end_time}},
        {"RadiusRatioA", 1, {&p->radius_ratio_a}},
        {"RadiusRatioP", 1, {&p->radius_ratio_p}},
        {"RadiusRatioV", 1, {&p->radius_ratio_v}},
        {"Density", 6, {&D[0], &D[1], &D[2], &D[3], &D[4], &D[5]}},
        {"BulkModulus", 6, {&K[0], &K[1], &K[2], &K[3], &K[4], &K[5]}},
        {"BulkViscosity", 6, {&L[0], &L[1], &L[2], &L[3], &L[4], &L[5]}},
        {"ShearViscosity", 6, {&M[0], &M[1], &M[2], &M[3], &M[4], &M[5]}},
        {"SurfaceTension", 4, {&S[0], &S[1], &S[4], &S[5]}},       // :756
        {"YoungModulus", 4, {&Y[2], &Y[3], &Y[4], &Y[5]}},         // :757
        {"PoissonRatio", 4, {&Po[2], &Po[3], &Po[4], &Po[5]}},     // :758
    };
    static const char *ir_names[6] = {"InteractionRatio(Type0)", "InteractionRatio(Type1)",
                                      "InteractionRatio(Type2)", "InteractionRatio(Type3)",
                                      "InteractionRatio(Type4)", "InteractionRatio(Type5)"};
    for (int t = 0; t < 6; ++t) {
        double *r = p->interaction_ratio[t];
        keys.push_back({ir_names[t], 6, {&r[0], &r[1], &r[2], &r[3], &r[4], &r[5]}});
    }
    keys.push_back({"Gravity", 3, {&p->gravity[0], &p->gravity[1], &p->gravity[2]}});

    char buf[1024];
    while (!std::feof(fp) && !std::ferror(fp)) {
        if (std::fgets(buf, sizeof(buf), fp) == NULL) continue;
        if (buf[0] == '#') continue;
        bool ok = false;
        for (size_t i = 0; i < keys.size() && !ok; ++i) ok = match_key(buf, keys[i]);
        if (!ok) ok = match_wall(buf, "Wall6", p, 4); // :766
        if (!ok) ok = match_wall(buf, "Wall7", p, 5); // :767
        if (!ok && on_invalid_line) on_invalid_line(buf, user); // :769
    }
    std::fclose(fp);
    return MPHX_OK;
}

extern "C" void mphx_free_host(void *ptr) { std::free(ptr); }

// The pre-processor's input (generator/generator.cpp:127-262: ParticleDistance, LowerDomain, UpperDomain and the
// StartCuboid ... EndCuboid blocks with Spacing / Type / Lower / Upper / Velocity; RigidType and Enthalpy are read and
// not used by the solver).  The header values take the same `%e` round trip the generator's .grid text gives them
// (generator.cpp:839-847), so a run from the .boid equals a run from the generated .grid.  Other shapes (Cuboid2,
// cylinders, rectangles) are not supported by the device-side generator.
static double through_e_text(double v)
{
    char b[64];
    std::snprintf(b, sizeof(b), "%e", v);
    return std::strtod(b, nullptr);
}
extern "C" int mphx_read_boid_file(const char *filename, mphx_params *p, mphx_cuboid **cuboids, int *ncuboids)
{
    if (!filename || !p || !cuboids || !ncuboids) return MPHX_ERR_INVALID;
    *cuboids = nullptr;
    *ncuboids = 0;
    FILE *fp = std::fopen(filename, "r");
    if (!fp) {
        mphx::set_last_error(std::string("cannot open boid file ") + filename);
        return MPHX_ERR_IO;
    }
    std::vector<mphx_cuboid> cubs;
    char buf[1024], token[256];
    bool in_cuboid = false, have_pd = false, have_lo = false, have_hi = false;
    unsigned seen = 0; // Spacing 1, Type 2, Lower 4, Upper 8
    mphx_cuboid cur{};
    int rc = MPHX_OK;
    double pd = 0.0, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    while (std::fgets(buf, sizeof(buf), fp)) {
        if (buf[0] == '#') continue;
        if (std::sscanf(buf, "%255s", token) != 1) continue;
        const std::string t = token;
        if (!in_cuboid) {
            if (t == "ParticleDistance") have_pd = std::sscanf(buf, " %*s %lf", &pd) == 1;
            else if (t == "LowerDomain") have_lo = std::sscanf(buf, " %*s %lf %lf %lf", &lo[0], &lo[1], &lo[2]) == 3;
            else if (t == "UpperDomain") have_hi = std::sscanf(buf, " %*s %lf %lf %lf", &hi[0], &hi[1], &hi[2]) == 3;
            else if (t == "StartCuboid") { in_cuboid = true; cur = mphx_cuboid{}; seen = 0; }
            else if (t.compare(0, 5, "Start") == 0) {
                mphx::set_last_error("boid file: shape '" + t + "' is not supported (cuboids only)");
                rc = MPHX_ERR_UNSUPPORTED;
                break;
            }
            continue;
        }
        if (t == "EndCuboid") {
            if (seen != 15u) { mphx::set_last_error("boid file: a cuboid lacks Spacing, Type, Lower or Upper"); rc = MPHX_ERR_IO; break; }
            cubs.push_back(cur);
            in_cuboid = false;
        } else if (t == "Spacing") { if (std::sscanf(buf, " %*s %lf", &cur.spacing) == 1) seen |= 1u; }
        else if (t == "Type") { if (std::sscanf(buf, " %*s %d", &cur.type) == 1) seen |= 2u; }
        else if (t == "Lower") { if (std::sscanf(buf, " %*s %lf %lf %lf", &cur.lower[0], &cur.lower[1], &cur.lower[2]) == 3) seen |= 4u; }
        else if (t == "Upper") { if (std::sscanf(buf, " %*s %lf %lf %lf", &cur.upper[0], &cur.upper[1], &cur.upper[2]) == 3) seen |= 8u; }
        else if (t == "Velocity") std::sscanf(buf, " %*s %lf %lf %lf", &cur.velocity[0], &cur.velocity[1], &cur.velocity[2]);
        // RigidType, Enthalpy: pre-processor fields the solver does not read
    }
    std::fclose(fp);
    if (rc) return rc;
    if (in_cuboid || !have_pd || !have_lo || !have_hi || cubs.empty()) {
        mphx::set_last_error("boid file: ParticleDistance, LowerDomain, UpperDomain and at least one complete cuboid are required");
        return MPHX_ERR_IO;
    }
    p->time0 = 0.0; // generator.cpp:841
    p->particle_spacing = through_e_text(pd);
    for (int d = 0; d < 3; ++d) { p->domain_min[d] = through_e_text(lo[d]); p->domain_max[d] = through_e_text(hi[d]); }
    mphx_cuboid *out = (mphx_cuboid *)std::calloc(cubs.size(), sizeof(mphx_cuboid));
    if (!out) return MPHX_ERR_NOMEM;
    std::memcpy(out, cubs.data(), sizeof(mphx_cuboid) * cubs.size());
    *cuboids = out;
    *ncuboids = (int)cubs.size();
    return MPHX_OK;
}

extern "C" void mphx_class_ranges(int n, const int *property, int ranges[6])
{
    for (int i = 0; i < 6; ++i) ranges[i] = -1;
    for (int i = 0; i < n; ++i) {
        const int t = property[i];
        const int cls = (0 <= t && t < 2) ? 0 : (2 <= t && t < 4) ? 1 : (4 <= t && t < 6) ? 2 : -1;
        if (cls < 0) continue;
        if (ranges[2 * cls] == -1) ranges[2 * cls] = i;
        ranges[2 * cls + 1] = i + 1;
    }
}

extern "C" int mphx_read_grid_file(const char *filename, mphx_params *p, int *n_out, int **property,
                                   double **position, double **initial_position, double **velocity)
{
    if (!filename || !p || !n_out || !property || !position || !initial_position || !velocity)
        return MPHX_ERR_INVALID;
    FILE *fp = std::fopen(filename, "r");
    if (!fp) {
        mphx::set_last_error(std::string("cannot open grid file ") + filename);
        return MPHX_ERR_IO;
    }
    char buf[1024];
    int n = 0;
    if (!std::fgets(buf, sizeof(buf), fp)) { std::fclose(fp); return MPHX_ERR_IO; }
    std::sscanf(buf, "%lf", &p->time0); // :797
    if (!std::fgets(buf, sizeof(buf), fp)) { std::fclose(fp); return MPHX_ERR_IO; }
    if (std::sscanf(buf, "%d  %lf  %lf %lf %lf  %lf %lf %lf", &n, &p->particle_spacing, // :799-804
                    &p->domain_min[0], &p->domain_max[0], &p->domain_min[1], &p->domain_max[1],
                    &p->domain_min[2], &p->domain_max[2]) != 8 || n < 0) {
        std::fclose(fp);
        mphx::set_last_error("malformed grid header");
        return MPHX_ERR_IO;
    }
    const size_t N = (size_t)n;
    int *t = (int *)std::calloc(N ? N : 1, sizeof(int));
    double *x = (double *)std::calloc(N ? 3 * N : 1, sizeof(double));
    double *x0 = (double *)std::calloc(N ? 3 * N : 1, sizeof(double));
    double *v = (double *)std::calloc(N ? 3 * N : 1, sizeof(double));
    if (!t || !x || !x0 || !v) {
        std::free(t); std::free(x); std::free(x0); std::free(v);
        std::fclose(fp);
        return MPHX_ERR_NOMEM;
    }
    for (size_t i = 0; i < N; ++i) { // :896-904; strtol/strtod parse exactly like %d/%lf
        if (!std::fgets(buf, sizeof(buf), fp)) break;
        char *s = buf, *e;
        t[i] = (int)std::strtol(s, &e, 10);
        if (e == s) continue;
        s = e;
        double vals[9];
        int k = 0;
        for (; k < 9; ++k) {
            vals[k] = std::strtod(s, &e);
            if (e == s) break;
            s = e;
        }
        for (int d = 0; d < 3; ++d) {
            if (d < k) x[3 * i + d] = vals[d];
            if (3 + d < k) x0[3 * i + d] = vals[3 + d];
            if (6 + d < k) v[3 * i + d] = vals[6 + d];
        }
    }
    std::fclose(fp);
    *n_out = n;
    *property = t;
    *position = x;
    *initial_position = x0;
    *velocity = v;
    return MPHX_OK;
}

namespace {
// big buffered writer: at 10M particles the per-line fprintf of the reference dominates wall time
struct Out {
    FILE *fp;
    std::vector<char> buf;
    size_t used = 0;
    explicit Out(FILE *f) : fp(f), buf(1 << 22) {}
    char *reserve(size_t n)
    {
        if (used + n > buf.size()) flush();
        return buf.data() + used;
    }
    void flush()
    {
        if (used) std::fwrite(buf.data(), 1, used, fp);
        used = 0;
    }
    void puts(const char *s)
    {
        size_t n = std::strlen(s);
        std::memcpy(reserve(n), s, n);
        used += n;
    }
};
} // namespace

extern "C" int mphx_write_prof_file(const char *filename, double time, const mphx_params *p, int n,
                                    const int *property, const double *position,
                                    const double *initial_position, const double *velocity)
{
    if (!filename || !p || n < 0) return MPHX_ERR_INVALID;
    FILE *fp = std::fopen(filename, "w");
    if (!fp) {
        mphx::set_last_error(std::string("cannot open prof file ") + filename);
        return MPHX_ERR_IO;
    }
    Out o(fp);
    o.used += std::snprintf(o.reserve(64), 64, "%e\n", time);
    o.used += std::snprintf(o.reserve(256), 256, "%d %e %e %e %e %e %e %e\n", n, p->particle_spacing,
                            p->domain_min[0], p->domain_max[0], p->domain_min[1], p->domain_max[1],
                            p->domain_min[2], p->domain_max[2]);
    for (int i = 0; i < n; ++i) {
        const double *x = position + 3 * (size_t)i, *x0 = initial_position + 3 * (size_t)i;
        const double *v = velocity + 3 * (size_t)i;
        o.used += std::snprintf(o.reserve(320), 320, "%d %e %e %e %e %e %e  %e %e %e\n", property[i],
                                x[0], x[1], x[2], x0[0], x0[1], x0[2], v[0], v[1], v[2]);
    }
    o.flush();
    std::fflush(fp);
    std::fclose(fp);
    return MPHX_OK;
}

extern "C" int mphx_write_vtk_file(const char *filename, int n, const double *initial_position,
                                   const mphx_host_views *f)
{
    if (!filename || !f || n < 0 || !initial_position) return MPHX_ERR_INVALID;
    if (!f->property || !f->position || !f->velocity || !f->force || !f->acceleration || !f->stress ||
        !f->strain || !f->neighbor_count || !f->initial_structure_neighbor_count)
        return MPHX_ERR_INVALID;
    FILE *fp = std::fopen(filename, "w");
    if (!fp) {
        mphx::set_last_error(std::string("cannot open vtk file ") + filename);
        return MPHX_ERR_IO;
    }
    Out o(fp);
    char *b;
    auto vec3f = [&](const double *a) { // "%e %e %e\n" of three float-cast values
        b = o.reserve(128);
        o.used += std::snprintf(b, 128, "%e %e %e\n", (float)a[0], (float)a[1], (float)a[2]);
    };
    o.puts("# vtk DataFile Version 2.0\n");
    o.puts("Unstructured Grid Example\n");
    o.puts("ASCII\n");
    o.puts("DATASET UNSTRUCTURED_GRID\n");
    o.used += std::snprintf(o.reserve(64), 64, "POINTS %d float\n", n);
    for (int i = 0; i < n; ++i) vec3f(f->position + 3 * (size_t)i);
    o.used += std::snprintf(o.reserve(64), 64, "CELLS %d %d\n", n, 2 * n);
    for (int i = 0; i < n; ++i) o.used += std::snprintf(o.reserve(32), 32, "1 %d ", i);
    o.puts("\n");
    o.used += std::snprintf(o.reserve(64), 64, "CELL_TYPES %d\n", n);
    for (int i = 0; i < n; ++i) o.puts("1 ");
    o.puts("\n");
    o.puts("\n");
    o.used += std::snprintf(o.reserve(64), 64, "POINT_DATA %d\n", n);
    o.puts("SCALARS label float 1\n");
    o.puts("LOOKUP_TABLE default\n");
    for (int i = 0; i < n; ++i) o.used += std::snprintf(o.reserve(32), 32, "%d\n", f->property[i]);
    o.puts("\n");
    o.puts("\n");
    o.puts("VECTORS displacement float\n");
    for (int i = 0; i < n; ++i) {
        const double *x = f->position + 3 * (size_t)i, *x0 = initial_position + 3 * (size_t)i;
        const double d[3] = {x[0] - x0[0], x[1] - x0[1], x[2] - x0[2]};
        vec3f(d);
    }
    for (int pass = 0; pass < 2; ++pass) { // stress00..22 then strain00..22  (:1030-1047)
        const double *T = pass == 0 ? f->stress : f->strain;
        const char *nm = pass == 0 ? "stress" : "strain";
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) {
                o.puts("\n");
                o.used += std::snprintf(o.reserve(64), 64, " SCALARS %s%d%d float \n", nm, a, c);
                o.puts("LOOKUP_TABLE default\n");
                for (int i = 0; i < n; ++i)
                    o.used += std::snprintf(o.reserve(32), 32, "%e\n", (float)T[9 * (size_t)i + 3 * a + c]);
            }
    }
    o.puts("VECTORS velocity float\n");
    for (int i = 0; i < n; ++i) vec3f(f->velocity + 3 * (size_t)i);
    o.puts("\n");
    o.puts("VECTORS accel float\n");
    for (int i = 0; i < n; ++i) vec3f(f->acceleration + 3 * (size_t)i);
    o.puts("\n");
    o.puts("SCALARS Initialneighbor float 1\n");
    o.puts("LOOKUP_TABLE default\n");
    for (int i = 0; i < n; ++i)
        o.used += std::snprintf(o.reserve(32), 32, "%d\n", f->initial_structure_neighbor_count[i]);
    o.puts("SCALARS neighbor float 1\n");
    o.puts("LOOKUP_TABLE default\n");
    for (int i = 0; i < n; ++i) o.used += std::snprintf(o.reserve(32), 32, "%d\n", f->neighbor_count[i]);
    // the virial sections the reference keeps commented out (:1128-1143), written when the caller asks for them
    if (f->virial_pressure) {
        o.puts("SCALARS VirialPressureAtParticle float 1\n");
        o.puts("LOOKUP_TABLE default\n");
        for (int i = 0; i < n; ++i) o.used += std::snprintf(o.reserve(32), 32, "%e\n", (float)f->virial_pressure[i]);
        o.puts("\n");
    }
    if (f->virial_stress)
        for (int a = 0; a < 2; ++a)      // (iD, jD < DIM-1 with DIM = 3: the four in-plane components, :1134-1135)
            for (int c = 0; c < 2; ++c) {
                o.used += std::snprintf(o.reserve(64), 64, "SCALARS VirialStressAtParticle[%d][%d] float 1\n", a, c);
                o.puts("LOOKUP_TABLE default\n");
                for (int i = 0; i < n; ++i) o.used += std::snprintf(o.reserve(32), 32, "%e\n", (float)f->virial_stress[9 * (size_t)i + 3 * a + c]);
                o.puts("\n");
            }
    o.puts("VECTORS velocity float\n"); // Q8: the section appears twice (:1062 and :1169)
    for (int i = 0; i < n; ++i) vec3f(f->velocity + 3 * (size_t)i);
    o.puts("\n");
    o.puts("VECTORS force float\n");
    for (int i = 0; i < n; ++i) vec3f(f->force + 3 * (size_t)i);
    o.puts("\n");
    o.flush();
    std::fflush(fp);
    std::fclose(fp);
    return MPHX_OK;
}

// ---- lossless binary checkpoint ------------------------------------------------------------
// The reference restarts from its .prof files, which carry 7 significant digits (`%e`, :973-978): a restarted run is a
// different trajectory.  This format keeps every double bit for bit, plus what the text file does not carry at all: the
// wall centres (advanced every step, :3066-3070).  Little-endian, fixed layout:
//   "MPHXCKP1" | int32 n | int32 dim | f64 time | f64 spacing | f64 domain_min[3] | f64 domain_max[3] |
//   f64 wall_center[6][3] | int32 property[n] | f64 position[n][3] | f64 initial_position[n][3] | f64 velocity[n][3]
static const char kCkpMagic[8] = {'M', 'P', 'H', 'X', 'C', 'K', 'P', '1'};

extern "C" int mphx_write_checkpoint(const char *filename, double time, const mphx_params *p, int n, const int *property,
                                     const double *position, const double *initial_position, const double *velocity)
{
    if (!filename || !p || n <= 0 || !property || !position || !initial_position || !velocity) return MPHX_ERR_INVALID;
    FILE *fp = std::fopen(filename, "wb");
    if (!fp) { mphx::set_last_error(std::string("error in open ") + filename); return MPHX_ERR_IO; }
    bool ok = std::fwrite(kCkpMagic, 1, 8, fp) == 8;
    const int hdr[2] = {n, p->dim};
    ok = ok && std::fwrite(hdr, sizeof(int), 2, fp) == 2;
    ok = ok && std::fwrite(&time, sizeof(double), 1, fp) == 1;
    ok = ok && std::fwrite(&p->particle_spacing, sizeof(double), 1, fp) == 1;
    ok = ok && std::fwrite(p->domain_min, sizeof(double), 3, fp) == 3 && std::fwrite(p->domain_max, sizeof(double), 3, fp) == 3;
    ok = ok && std::fwrite(p->wall_center, sizeof(double), 3 * MPHX_TYPE_COUNT, fp) == 3 * MPHX_TYPE_COUNT;
    const size_t N = (size_t)n;
    ok = ok && std::fwrite(property, sizeof(int), N, fp) == N && std::fwrite(position, sizeof(double), 3 * N, fp) == 3 * N &&
         std::fwrite(initial_position, sizeof(double), 3 * N, fp) == 3 * N && std::fwrite(velocity, sizeof(double), 3 * N, fp) == 3 * N;
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok) { mphx::set_last_error(std::string("error writing ") + filename); return MPHX_ERR_IO; }
    return MPHX_OK;
}

extern "C" int mphx_read_checkpoint(const char *filename, mphx_params *p, int *n_out, int **property, double **position,
                                    double **initial_position, double **velocity)
{
    if (!filename || !p || !n_out || !property || !position || !initial_position || !velocity) return MPHX_ERR_INVALID;
    *property = nullptr; *position = *initial_position = *velocity = nullptr;
    FILE *fp = std::fopen(filename, "rb");
    if (!fp) { mphx::set_last_error(std::string("error in open ") + filename); return MPHX_ERR_IO; }
    char magic[8];
    int hdr[2] = {0, 0};
    double time = 0.0;
    bool ok = std::fread(magic, 1, 8, fp) == 8 && !std::memcmp(magic, kCkpMagic, 8) && std::fread(hdr, sizeof(int), 2, fp) == 2 && hdr[0] > 0 &&
              std::fread(&time, sizeof(double), 1, fp) == 1 && std::fread(&p->particle_spacing, sizeof(double), 1, fp) == 1 &&
              std::fread(p->domain_min, sizeof(double), 3, fp) == 3 && std::fread(p->domain_max, sizeof(double), 3, fp) == 3 &&
              std::fread(p->wall_center, sizeof(double), 3 * MPHX_TYPE_COUNT, fp) == 3 * MPHX_TYPE_COUNT;
    if (!ok) { std::fclose(fp); mphx::set_last_error(std::string(filename) + " is not an mphx checkpoint"); return MPHX_ERR_IO; }
    const size_t N = (size_t)hdr[0];
    int *t = (int *)std::malloc(sizeof(int) * N);
    double *x = (double *)std::malloc(sizeof(double) * 3 * N), *x0 = (double *)std::malloc(sizeof(double) * 3 * N),
           *v = (double *)std::malloc(sizeof(double) * 3 * N);
    if (!t || !x || !x0 || !v) { std::free(t); std::free(x); std::free(x0); std::free(v); std::fclose(fp); return MPHX_ERR_NOMEM; }
    ok = std::fread(t, sizeof(int), N, fp) == N && std::fread(x, sizeof(double), 3 * N, fp) == 3 * N &&
         std::fread(x0, sizeof(double), 3 * N, fp) == 3 * N && std::fread(v, sizeof(double), 3 * N, fp) == 3 * N;
    std::fclose(fp);
    if (!ok) { std::free(t); std::free(x); std::free(x0); std::free(v); mphx::set_last_error(std::string(filename) + " is truncated"); return MPHX_ERR_IO; }
    p->time0 = time;
    p->dim = hdr[1];
    *n_out = hdr[0]; *property = t; *position = x; *initial_position = x0; *velocity = v;
    return MPHX_OK;
}
