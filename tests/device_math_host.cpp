// TEST INFRASTRUCTURE.  The device-side limb arithmetic of the kernels (qimcifa_b200/csrc/qcf_device.cuh) compiled for
// the HOST with g++ -- the same source, not a restatement -- behind a C interface that tests/test_device_math.py
// drives with ctypes and checks against Python integers.  Built into tests/_build/libdevice_math_host.so by the test.
#include "qcf_device.cuh"
#include "qcf_filter_math.cuh"

namespace {
template <int L> void load(u32 (&d)[L], const u32* s)
{
    for (int i = 0; i < L; ++i) {
        d[i] = s[i];
    }
}

template <int LN, int LG, int QN> int divides_t(const u32* n, const u32* g, int form, u32 ninv)
{
    u32 nn[LN], gg[LG];
    load<LN>(nn, n);
    load<LG>(gg, g);
    switch (form) {
    case 0:
        return qcf_divides<LN, LG, QN>(nn, gg) ? 1 : 0;
    case 1:
        return qcf_divides_ninv<LN, LG, QN>(nn, gg, ninv) ? 1 : 0;
    default:
        return -1;
    }
}
template <int LN, int LG, int QN> int lowlimb_t(const u32* n, const u32* g, u32 ninv)
{
    u32 nn[LN], gg[LG];
    load<LN>(nn, n);
    load<LG>(gg, g);
    return qcf_lowlimb<LN, LG, QN>(nn, gg, ninv) ? 1 : 0;
}
template <int LN, int QN> int divides2_t(const u32* n, const u32* g, u32 ninv)
{
    u32 nn[LN];
    load<LN>(nn, n);
    return qcf_divides_ninv2<LN, QN>(nn, g[0], g[1], ninv) ? 1 : 0;
}
} // namespace

namespace {
// The chain walk of qcf_filter_kernel with the kernel's own arithmetic (qcf_filter_math.cuh): chain setup, trip offsets,
// per-segment re-basing, trip advance.  out[k] = the value the kernel compares for step k, for every step walked.
template <int D, u32 U> void filter_chain_t(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg,
    const u32* seg_trips, const u32* seg_shift, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c, u32 chain_r, u32 pinv32,
    u64 sigma0, const u32* seg_kappa, u32* out)
{
    u32 z, d1, elo, d2, d3;
    const u32 rho32 = rho_on ? qcf_filter_rho32(chain_c, chain_r, pinv32, tp) : 0u;
    qcf_filter_chain_init<D>(n64, g0, P, tp, al, c_lo, mu0, rho_on ? qcf_filter_rho_corr(rho32, sigma0) : 0ull, slack, z, d1, elo, d2, d3);
    u32 s[U];
    qcf_filter_offsets<U>(s, d1, d2, d3);
    const u32 f3 = U * d3;
    u64 k = 0;
    for (u32 sg = 0; sg < n_seg; ++sg) {
        if (sg == 0) {
            z += seg_shift[0];
        } else {
            qcf_filter_enter_segment<U>(z, d1, elo, s, seg_shift[sg], seg_nu_lo[sg], seg_nu_hi[sg], rho32, seg_kappa[sg]);
        }
        for (u32 t = 0; t < seg_trips[sg]; ++t) {
            for (u32 u = 0; u < U; ++u) {
                out[k++] = z + s[u];
            }
            qcf_filter_advance<D, U>(z, d1, d2, d3, f3, s);
        }
    }
}

// The degree-1 walk as the kernel makes it: packed 16-bit trips (qcf_filter_pack_offsets / qcf_filter_trip16), the packed
// offsets rebuilt at every segment.  out[k] = z + u*d1, the value the full-width replay compares; flag[t] = 1 when trip t's
// packed test asks for that replay.
template <u32 U> void filter_chain16_t(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg,
    const u32* seg_trips, const u32* seg_shift, const u32* seg_bound, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c,
    u32 chain_r, u32 pinv32, u64 sigma0, const u32* seg_kappa, u32* out, u32* flag)
{
    u32 z, d1, elo, d2, d3;
    const u32 rho32 = rho_on ? qcf_filter_rho32(chain_c, chain_r, pinv32, tp) : 0u;
    qcf_filter_chain_init<1>(n64, g0, P, tp, al, c_lo, mu0, rho_on ? qcf_filter_rho_corr(rho32, sigma0) : 0ull, slack, z, d1, elo, d2, d3);
    u32 S2[U / 2];
    qcf_filter_pack_offsets<U>(S2, d1);
    u64 k = 0, t_all = 0;
    for (u32 sg = 0; sg < n_seg; ++sg) {
        if (sg == 0) {
            z += seg_shift[0];
        } else {
            qcf_filter_enter_segment1(z, d1, elo, seg_shift[sg], seg_nu_lo[sg], seg_nu_hi[sg], rho32, seg_kappa[sg]);
            qcf_filter_pack_offsets<U>(S2, d1);
        }
        const u32 t2 = ((seg_bound[sg] >> 16) + 2u) * 0x10001u;
        for (u32 t = 0; t < seg_trips[sg]; ++t) {
            u32 m0 = t2, m1 = t2;
            flag[t_all++] = qcf_filter_trip16<U>(z, S2, m0, m1, t2) ? qcf_filter_flag_classes(m0, m1, t2) : 0u; // the classes (u mod 4) the replay looks at
            for (u32 u = 0; u < U; ++u) {
                out[k++] = z + u * d1;
            }
            z += U * d1;
        }
    }
}
// The degree-2 walk in the packed form (sub-trips of 16 in groups of 8: qcf_filter_pack2_*, qcf_filter_subtrip16): whole groups
// from the exact base, the rest of a run with the sub-trip index counted up and the base moved on by packed adds, the base
// rebuilt at every segment -- the control flow of qcf_filter_kernel<2, ., ., true> for runs of `run_trips` trips (an epoch
// budget that cuts the segments).  out[k] = the full-width value of step k (what the replay compares), flag[t] per trip.
void filter_chain_pk2(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg, const u32* seg_trips,
    const u32* seg_shift, const u32* seg_bound, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c, u32 chain_r, u32 pinv32,
    u64 sigma0, const u32* seg_kappa, u32 run_trips, u32* out, u32* flag)
{
    constexpr u32 U = 2u * QCF_FILTER_SUB, TG = QCF_FILTER_GROUP / 2u;
    u32 z, d1, elo, d2, d3;
    const u32 rho32 = rho_on ? qcf_filter_rho32(chain_c, chain_r, pinv32, tp) : 0u;
    qcf_filter_chain_init<2>(n64, g0, P, tp, al, c_lo, mu0, rho_on ? qcf_filter_rho_corr(rho32, sigma0) : 0ull, slack, z, d1, elo, d2, d3);
    u32 S2[QCF_FILTER_SUB / 2], E2[QCF_FILTER_SUB / 2], EG[QCF_FILTER_SUB / 2];
    qcf_filter_pack2_moves(E2, EG, d2);
    qcf_filter_pack2_base(S2, E2, d1, d2);
    u64 k = 0, t_all = 0;
    u32 budget = run_trips;
    u32 a = qcf_filter_sub_move(d1, d2), d1z = d1; // as the kernel: the limb's move per sub-trip; d1 at step k = d1z + k * d2
    for (u32 sg = 0; sg < n_seg; ++sg) {
        if (sg == 0) {
            z += seg_shift[0];
        } else {
            const u32 d1k = d1z + (u32)k * d2;
            u32 d1n = d1k;
            qcf_filter_enter_segment1(z, d1n, elo, seg_shift[sg], seg_nu_lo[sg], seg_nu_hi[sg], rho32, seg_kappa[sg]);
            d1z += d1n - d1k;
            a += QCF_FILTER_SUB * (d1n - d1k);
            qcf_filter_pack2_base(S2, E2, d1n, d2);
        }
        const u32 t2 = qcf_filter_t2_group(seg_bound[sg]);
        const auto trip = [&](const u32 j) {
            u32 m0 = t2, m1 = t2;
            u32 v = z, dv = d1z + (u32)k * d2; // what the replay walks from: the first difference at the trip's first step
            for (u32 u = 0; u < U; ++u) {
                out[k++] = v;
                v += dv;
                dv += d2;
            }
            qcf_filter_subtrip16(z, j, S2, E2, m0, m1);
            qcf_filter_sub_advance(z, a, d2);
            qcf_filter_subtrip16(z, j + 1u, S2, E2, m0, m1);
            qcf_filter_sub_advance(z, a, d2);
            flag[t_all++] = qcf_filter_flag_classes(m0, m1, t2); // 0 = not flagged, else the classes (u mod 4) the replay looks at
        };
        for (u32 left = seg_trips[sg]; left > 0u;) {
            const u32 run = (left < budget) ? left : budget;
            left -= run;
            budget -= run;
            u32 r = run;
            for (; r >= TG; r -= TG) {
                for (u32 t = 0; t < TG; ++t) {
                    trip(2u * t);
                }
                for (u32 p = 0; p < QCF_FILTER_SUB / 2; ++p) {
                    S2[p] = qcf_add16x2(S2[p], EG[p]);
                }
            }
            if (r) {
                u32 j = 0;
                for (; r > 0u; --r, j += 2u) {
                    trip(j);
                }
                if (left > 0u) { // as the kernel: at the end of a segment the next one rebuilds the base
                    for (; j > 0u; --j) {
                        qcf_filter_pack2_step_base(S2, E2);
                    }
                }
            }
            if (budget == 0u) {
                budget = run_trips;
            }
        }
    }
}
} // namespace

extern "C" {

u32 h_filter_chain_pk2(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg, const u32* seg_trips,
    const u32* seg_shift, const u32* seg_bound, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c, u32 chain_r, u32 pinv32,
    u64 sigma0, const u32* seg_kappa, u32 run_trips, u32* out, u32* flag)
{
    filter_chain_pk2(n64, g0, P, tp, al, c_lo, mu0, slack, n_seg, seg_trips, seg_shift, seg_bound, seg_nu_lo, seg_nu_hi, rho_on, chain_c,
        chain_r, pinv32, sigma0, seg_kappa, run_trips, out, flag);
    return 2u * QCF_FILTER_SUB;
}

// min over the steps u = r (mod 4) of a trip of U steps, as the replay of a flagged packed trip computes it; cls = class mask
u32 h_filter_flagged_min(u32 U, u32 z, u32 d1, u32 d2, u32 cls)
{
    return (U == 128u) ? qcf_filter_flagged_min<128>(z, d1, d2, cls)
        : (U == 64u)   ? qcf_filter_flagged_min<64>(z, d1, d2, cls)
        : (U == 32u)   ? qcf_filter_flagged_min<32>(z, d1, d2, cls)
                       : qcf_filter_flagged_min<16>(z, d1, d2, cls);
}

u32 h_filter_chain16(u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg, const u32* seg_trips,
    const u32* seg_shift, const u32* seg_bound, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c, u32 chain_r, u32 pinv32,
    u64 sigma0, const u32* seg_kappa, u32* out, u32* flag)
{
    filter_chain16_t<QCF_FILTER_UNROLL_FOR(1)>(n64, g0, P, tp, al, c_lo, mu0, slack, n_seg, seg_trips, seg_shift, seg_bound, seg_nu_lo, seg_nu_hi,
        rho_on, chain_c, chain_r, pinv32, sigma0, seg_kappa, out, flag);
    return QCF_FILTER_UNROLL_FOR(1);
}

// degree 1..3; the trip length is the kernel's (32 guesses at degree <= 2, 16 at degree 3); returns it, 0 = bad degree
u32 h_filter_chain(int degree, u64 n64, u64 g0, u32 P, u32 tp, u32 al, u64 c_lo, u64 mu0, u32 slack, u32 n_seg, const u32* seg_trips,
    const u32* seg_shift, const u32* seg_nu_lo, const u32* seg_nu_hi, u32 rho_on, u64 chain_c, u32 chain_r, u32 pinv32, u64 sigma0,
    const u32* seg_kappa, u32* out)
{
    switch (degree) {
    case 1:
        filter_chain_t<1, QCF_FILTER_UNROLL_FOR(1)>(n64, g0, P, tp, al, c_lo, mu0, slack, n_seg, seg_trips, seg_shift, seg_nu_lo, seg_nu_hi, rho_on, chain_c, chain_r, pinv32, sigma0, seg_kappa, out);
        return QCF_FILTER_UNROLL_FOR(1);
    case 2:
        filter_chain_t<2, QCF_FILTER_UNROLL_FOR(2)>(n64, g0, P, tp, al, c_lo, mu0, slack, n_seg, seg_trips, seg_shift, seg_nu_lo, seg_nu_hi, rho_on, chain_c, chain_r, pinv32, sigma0, seg_kappa, out);
        return QCF_FILTER_UNROLL_FOR(2);
    case 3:
        filter_chain_t<3, QCF_FILTER_UNROLL_FOR(3)>(n64, g0, P, tp, al, c_lo, mu0, slack, n_seg, seg_trips, seg_shift, seg_nu_lo, seg_nu_hi, rho_on, chain_c, chain_r, pinv32, sigma0, seg_kappa, out);
        return QCF_FILTER_UNROLL_FOR(3);
    default:
        return 0;
    }
}

u32 h_inv32(u32 g) { return qcf_inv32(g); }
u32 h_ninv_step(u32 y, u32 g0) { return qcf_ninv_step(y, g0); }

// form 0: qcf_divides (reciprocal from scratch); form 1: qcf_divides_ninv; form 2: qcf_divides_ninv2 (LG == 2);
// form 3: qcf_divides_q2g2 (LG == 2, QN == 2, LN >= 3); form 4: qcf_lowlimb_q2g2 (stage 1 only); form 5: qcf_lowlimb (stage 1,
// any shape).  -1: no such shape.
int h_divides(int form, int ln, int lg, int qn, const u32* n, const u32* g, u32 ninv)
{
#define SHAPE(LN, LG, QN)                                                                                              \
    if ((ln == LN) && (lg == LG) && (qn == QN)) {                                                                      \
        if (form <= 1) {                                                                                               \
            return divides_t<LN, LG, QN>(n, g, form, ninv);                                                            \
        }                                                                                                              \
    }
    SHAPE(2, 1, 1) SHAPE(2, 1, 2) SHAPE(2, 2, 1) SHAPE(2, 2, 2)
    SHAPE(3, 1, 2) SHAPE(3, 1, 3) SHAPE(3, 2, 1) SHAPE(3, 2, 2) SHAPE(3, 2, 3)
    SHAPE(4, 1, 3) SHAPE(4, 1, 4) SHAPE(4, 2, 2) SHAPE(4, 2, 3) SHAPE(4, 2, 4) SHAPE(4, 3, 1) SHAPE(4, 3, 2)
    SHAPE(2, 2, 2) SHAPE(3, 3, 3) SHAPE(4, 3, 4) SHAPE(4, 2, 4)
#undef SHAPE
#define SHAPE2(LN, QN)                                                                                                 \
    if ((form == 2) && (lg == 2) && (ln == LN) && (qn == QN)) {                                                        \
        return divides2_t<LN, QN>(n, g, ninv);                                                                         \
    }
    SHAPE2(2, 1) SHAPE2(2, 2) SHAPE2(3, 1) SHAPE2(3, 2) SHAPE2(3, 3) SHAPE2(4, 2) SHAPE2(4, 3) SHAPE2(4, 4)
#undef SHAPE2
#define SHAPE5(LN, LG, QN)                                                                                             \
    if ((form == 5) && (ln == LN) && (lg == LG) && (qn == QN)) {                                                       \
        return lowlimb_t<LN, LG, QN>(n, g, ninv);                                                                      \
    }
    SHAPE5(2, 1, 1) SHAPE5(2, 1, 2) SHAPE5(2, 2, 1) SHAPE5(2, 2, 2) SHAPE5(3, 1, 2) SHAPE5(3, 1, 3) SHAPE5(3, 2, 1) SHAPE5(3, 2, 2)
    SHAPE5(3, 2, 3) SHAPE5(4, 1, 3) SHAPE5(4, 1, 4) SHAPE5(4, 2, 2) SHAPE5(4, 2, 3) SHAPE5(4, 2, 4) SHAPE5(4, 3, 1) SHAPE5(4, 3, 2)
    SHAPE5(3, 3, 3) SHAPE5(4, 3, 4)
#undef SHAPE5
    if ((form == 3 || form == 4) && (lg == 2) && (qn == 2) && (ln == 3 || ln == 4)) {
        u32 n3[3], n4[4];
        load<3>(n3, n);
        load<4>(n4, n);
        if (form == 3) {
            return ((ln == 3) ? qcf_divides_q2g2<3>(n3, g[0], g[1], ninv) : qcf_divides_q2g2<4>(n4, g[0], g[1], ninv)) ? 1 : 0;
        }
        return ((ln == 3) ? qcf_lowlimb_q2g2<3>(n3, g[0], g[1], ninv) : qcf_lowlimb_q2g2<4>(n4, g[0], g[1], ninv)) ? 1 : 0;
    }
    return -1;
}

// a += b on lg limbs (the host build takes the portable branch of qcf_addn; the carry-flag PTX is device only)
void h_addn(int lg, u32* a, const u32* b)
{
    if (lg == 1) {
        u32 x[1] = { a[0] }, y[1] = { b[0] };
        qcf_addn<1>(x, y);
        a[0] = x[0];
    } else if (lg == 2) {
        u32 x[2] = { a[0], a[1] }, y[2] = { b[0], b[1] };
        qcf_addn<2>(x, y);
        a[0] = x[0], a[1] = x[1];
    } else {
        u32 x[3] = { a[0], a[1], a[2] }, y[3] = { b[0], b[1], b[2] };
        qcf_addn<3>(x, y);
        a[0] = x[0], a[1] = x[1], a[2] = x[2];
    }
}

void h_add32(u32* r, const u32* a, u32 b)
{
    u32 x[3] = { a[0], a[1], a[2] }, y[3];
    qcf_add32<3>(y, x, b);
    r[0] = y[0], r[1] = y[1], r[2] = y[2];
}

int h_less(const u32* a, const u32* b)
{
    u32 x[3] = { a[0], a[1], a[2] }, y[3] = { b[0], b[1], b[2] };
    return qcf_less<3>(x, y) ? 1 : 0;
}
}
