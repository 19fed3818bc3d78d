/* Planner regression harness (host only, no GPU): loads two builds of libqimcifa_cuda.so and checks that qcf_plan_filter
 * returns bit-identical plans on random inputs, and times both.
 *   gcc -O2 -I include tools/plan_equiv.c -o /tmp/plan_equiv -ldl && /tmp/plan_equiv OLD.so NEW.so [cases]           */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "qimcifa_cuda.h"

typedef int (*plan_fn)(const uint32_t*, int, int, const uint32_t*, const uint32_t*, int, qcf_filter_plan*);
typedef unsigned __int128 u128;

static uint64_t s[2] = { 0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull };
static uint64_t rnd(void)
{ /* xorshift128+ */
    uint64_t a = s[0], b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static u128 rnd128(void) { return ((u128)rnd() << 64) | rnd(); }
static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}
static void limbs(u128 v, uint32_t* le)
{
    for (int i = 0; i < 4; i++) {
        le[i] = (uint32_t)v;
        v >>= 32;
    }
}
static u128 isqrt(u128 n)
{
    u128 x = (u128)1 << 64;
    for (;;) {
        u128 y = (x + n / x) >> 1;
        if (y >= x) {
            return x;
        }
        x = y;
    }
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s OLD.so NEW.so [cases]\n", argv[0]);
        return 2;
    }
    void* ho = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    void* hn = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!ho || !hn) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    plan_fn fo = (plan_fn)dlsym(ho, "qcf_plan_filter"), fn = (plan_fn)dlsym(hn, "qcf_plan_filter");
    const long cases = (argc > 3) ? atol(argv[3]) : 200000;
    long found = 0, diff = 0;
    double to = 0, tn = 0;
    static const int bits_choices[] = { 40, 56, 64, 72, 80, 96, 100, 112, 128 };
    static const char* forces[] = { NULL, NULL, NULL, NULL, "1,0,0,0", "2,0,0,0", "3,0,0,0", "0,0,32,0", "0,0,4,1", "0,0,8,2", "2,14,0,0", "0,0,1,0" };
    for (long c = 0; c < cases; c++) {
        const int bits = bits_choices[rnd() % 9];
        u128 N = (rnd128() >> (128 - bits)) | ((u128)1 << (bits - 1)) | 1;
        const u128 top = isqrt(N) >> (rnd() % 9);
        u128 v_hi = top > 1000000 ? top - (rnd() % 1000) : 1000000;
        const int sh = 18 + (int)(rnd() % 24);
        u128 span = ((u128)(3 + rnd() % 40)) << sh;
        u128 v_lo = (v_hi > span) ? v_hi - span : 17;
        if ((rnd() & 3) && (v_lo < v_hi / 2 + 1)) {
            v_lo = v_hi / 2 + 1; /* one octave at most, as run_values cuts it */
        }
        const int level = 5 + (int)(rnd() % 5), forced = (int)(rnd() & 1);
        const char* f = forces[rnd() % 12];
        if (f) {
            setenv("QCF_PLAN_FORCE", f, 1);
        } else {
            unsetenv("QCF_PLAN_FORCE");
        }
        uint32_t n[4], lo[4], hi[4];
        limbs(N, n);
        limbs(v_lo, lo);
        limbs(v_hi, hi);
        qcf_filter_plan po, pn;
        memset(&po, 0, sizeof(po));
        memset(&pn, 0, sizeof(pn));
        double t0 = now();
        const int ro = fo(n, 4, level, lo, hi, forced, &po);
        double t1 = now();
        const int rn = fn(n, 4, level, lo, hi, forced, &pn);
        double t2 = now();
        to += t1 - t0;
        tn += t2 - t1;
        found += po.found;
        if ((ro != rn) || memcmp(&po, &pn, sizeof(po))) {
            if (diff < 5) {
                fprintf(stderr, "DIFF case %ld bits %d level %d forced %d force %s: old found %d D %d tp %u K %u J %u cost %.6f | new found %d D %d tp %u K %u J %u cost %.6f\n",
                    c, bits, level, forced, f ? f : "-", po.found, po.degree, po.tprime, po.steps, po.n_seg, po.cost, pn.found, pn.degree,
                    pn.tprime, pn.steps, pn.n_seg, pn.cost);
            }
            diff++;
        }
    }
    printf("cases %ld, plans found %ld, differing %ld; old %.2f us/plan, new %.2f us/plan\n", cases, found, diff, to / cases * 1e6, tn / cases * 1e6);
    return diff ? 1 : 0;
}
