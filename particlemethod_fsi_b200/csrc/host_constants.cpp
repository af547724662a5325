// host_constants.cpp -- one-time host constants of the explicit step (no GPU needed).
//
// Replaces the reference's initializeWeight (src/main.cpp:1191-1309), the type-level part of
// initializeFluid (:1329-1341), initializeWall (:1371-1410) and initializeDomain (:1412-1469).
// These few scalars feed every kernel, so they are evaluated with the reference's operand order
// (no FMA: this TU is built with -ffp-contract=off) and come out bit-identical; the parity test
// compares them with == against the reference.
#include <cmath>
#include <cstring>

#include "mphx.h"
#include "mphx_internal.h"

namespace {

struct KernelShape {
    bool two_d;
    double hd(double h) const { return two_d ? h * h : h * h * h; }
    // wa (:299-305) and wp (:335-341) at lattice distance r
    double wa(double swa, double r, double h) const
    {
        return 1.0 / swa * 1.0 / hd(h) * (r / h) * (1.0 - (r / h)) * (1.0 - (r / h));
    }
    double wp(double swp, double r, double h) const
    {
        return 1.0 / swp * 1.0 / hd(h) * ((1.0 - r / h) * (1.0 - r / h));
    }
};

// N0a / N0p: sum of the kernel over a perfect lattice (:1216-1304), x outer, y, z inner
double lattice_reference_density(const KernelShape &ks, double l0, double radius, double sw,
                                 bool attractive, int *count)
{
    const int range = (int)(radius / l0 + 3.0);
    const int zr = ks.two_d ? 0 : range;
    double sum = 0.0;
    int cnt = 0;
    for (int ix = -range; ix <= range; ++ix)
        for (int iy = -range; iy <= range; ++iy)
            for (int iz = -zr; iz <= zr; ++iz) {
                if (ix == 0 && iy == 0 && iz == 0) continue;
                const double x = l0 * (double)ix, y = l0 * (double)iy, z = l0 * (double)iz;
                const double r2 = ks.two_d ? (x * x + y * y) : (x * x + y * y + z * z);
                if (r2 <= radius * radius) {
                    const double r = std::sqrt(r2);
                    sum += attractive ? ks.wa(sw, r, radius) : ks.wp(sw, r, radius);
                    ++cnt;
                }
            }
    *count = cnt;
    return sum;
}

} // namespace

extern "C" int mphx_compute_constants(const mphx_params *p, mphx_constants *c)
{
    if (!p || !c) return MPHX_ERR_INVALID;
    if (p->dim != 2 && p->dim != 3) return MPHX_ERR_INVALID;
    if (!(p->particle_spacing > 0.0)) return MPHX_ERR_INVALID;
    std::memset(c, 0, sizeof(*c));
    const bool two_d = (p->dim == 2);
    const KernelShape ks{two_d};
    const double l0 = p->particle_spacing;
    const double pi = M_PI;

    c->particle_volume = two_d ? l0 * l0 : l0 * l0 * l0; // :806-808
    c->radius_a = p->radius_ratio_a * l0;                // :1195-1198 (RadiusRatioG = RadiusRatioA)
    c->radius_g = p->radius_ratio_a * l0;
    c->radius_p = p->radius_ratio_p * l0;
    c->radius_v = p->radius_ratio_v * l0;
    if (two_d) { // :1202-1206
        c->swa = 1.0 / 2.0 * 2.0 / 15.0 * pi / l0 / l0;
        c->swg = 1.0 / 2.0 * 1.0 / 3.0 * pi / l0 / l0;
        c->swp = c->swv = c->swg;
        c->r2g = 1.0 / 2.0 * 1.0 / 30.0 * pi * c->radius_g * c->radius_g / l0 / l0 / c->swg;
    } else { // :1208-1212
        c->swa = 1.0 / 3.0 * 1.0 / 5.0 * pi / l0 / l0 / l0;
        c->swg = 1.0 / 3.0 * 2.0 / 5.0 * pi / l0 / l0 / l0;
        c->swp = c->swv = c->swg;
        c->r2g = 1.0 / 3.0 * 4.0 / 105.0 * pi * c->radius_g * c->radius_g / l0 / l0 / l0 / c->swg;
    }
    c->n0a = lattice_reference_density(ks, l0, c->radius_a, c->swa, true, &c->n0a_count);
    c->n0p = lattice_reference_density(ks, l0, c->radius_p, c->swp, false, &c->n0p_count);

    // surface-tension coefficients :1329-1341
    double integ_n, integ_x;
    if (two_d) { c->cof_k = 0.350778153; integ_n = 0.024679383; integ_x = 0.226126699; }
    else       { c->cof_k = 0.326976006; integ_n = 0.021425779; integ_x = 0.233977488; }
    for (int t = 0; t < MPHX_TYPE_COUNT; ++t)
        c->cof_a[t] = p->surface_tension[t] / ((c->radius_g / l0) * (integ_n + c->cof_k * c->cof_k * integ_x));

    // wall rotation per step, quaternion form :1374-1408.  Q9: the reference uses theta=|omega|^2.
    for (int t = 4; t < MPHX_TYPE_COUNT; ++t) {
        const double *w = p->wall_omega[t];
        const double theta = std::fabs(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        double nrm[3] = {0.0, 0.0, 0.0};
        if (theta != 0.0)
            for (int d = 0; d < 3; ++d) nrm[d] = w[d] / theta;
        const double s = std::sin(theta * p->dt / 2.0), q3 = std::cos(theta * p->dt / 2.0);
        const double q0 = nrm[0] * s, q1 = nrm[1] * s, q2 = nrm[2] * s;
        double(*R)[3] = c->wall_rotation[t];
        R[0][0] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
        R[0][1] = 2.0 * (q0 * q1 - q2 * q3);
        R[0][2] = 2.0 * (q0 * q2 + q1 * q3);
        R[1][0] = 2.0 * (q0 * q1 + q2 * q3);
        R[1][1] = -q0 * q0 + q1 * q1 - q2 * q2 + q3 * q3;
        R[1][2] = 2.0 * (q1 * q2 - q0 * q3);
        R[2][0] = 2.0 * (q0 * q2 - q1 * q3);
        R[2][1] = 2.0 * (q1 * q2 + q0 * q3);
        R[2][2] = -q0 * q0 - q1 * q1 + q2 * q2 + q3 * q3;
    }

    // background cells :1414-1440 (cell width = one particle spacing)
    c->cell_width = l0;
    double cc[3];
    cc[0] = std::round((p->domain_max[0] - p->domain_min[0]) / c->cell_width);
    cc[1] = std::round((p->domain_max[1] - p->domain_min[1]) / c->cell_width);
    cc[2] = two_d ? 1.0 : std::round((p->domain_max[2] - p->domain_min[2]) / c->cell_width);
    for (int d = 0; d < 3; ++d) {
        if (!(cc[d] >= 1.0) || cc[d] > 2.0e9) return MPHX_ERR_UNSUPPORTED;
        c->cell_count[d] = (int)cc[d];
        c->domain_max[d] = p->domain_max[d]; // the reference's fix-up branch (:1431) can never fire
        c->domain_width[d] = c->domain_max[d] - p->domain_min[d];
    }
    const double total = cc[0] * cc[1] * cc[2];
    if (total > 2.0e9) return MPHX_ERR_UNSUPPORTED; // `int CellCounts` (:1429) would overflow
    c->cell_counts = (int)total;

    // MaxRadius :1460-1463, list cut-off range :1744
    double mr = 0.0;
    mr = (c->radius_a > mr) ? c->radius_a : mr;
    mr = (c->radius_g > mr) ? c->radius_g : mr;
    mr = (c->radius_p > mr) ? c->radius_p : mr;
    mr = (c->radius_v > mr) ? c->radius_v : mr;
    c->max_radius = mr;
    c->stencil_range = (int)(std::ceil((mr + 0.1 * l0) / c->cell_width));
    return MPHX_OK;
}
