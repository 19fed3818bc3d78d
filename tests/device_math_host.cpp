// TEST INFRASTRUCTURE.  The device-side limb arithmetic of the kernels (qimcifa_b200/csrc/qcf_device.cuh) compiled for
// the HOST with g++ -- the same source, not a restatement -- behind a C interface that tests/test_device_math.py
// drives with ctypes and checks against Python integers.  Built into tests/_build/libdevice_math_host.so by the test.
#include "qcf_device.cuh"

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
template <int LN, int QN> int divides2_t(const u32* n, const u32* g, u32 ninv)
{
    u32 nn[LN];
    load<LN>(nn, n);
    return qcf_divides_ninv2<LN, QN>(nn, g[0], g[1], ninv) ? 1 : 0;
}
} // namespace

extern "C" {

u32 h_inv32(u32 g) { return qcf_inv32(g); }
u32 h_ninv_step(u32 y, u32 g0) { return qcf_ninv_step(y, g0); }

// form 0: qcf_divides (reciprocal from scratch); form 1: qcf_divides_ninv; form 2: qcf_divides_ninv2 (LG == 2);
// form 3: qcf_divides_q2g2 (LG == 2, QN == 2, LN >= 3); form 4: qcf_lowlimb_q2g2 (stage 1 only).  -1: no such shape.
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
