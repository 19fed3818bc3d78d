// Compat stand-in for <boost/multiprecision/cpp_int.hpp> (TEST INFRASTRUCTURE ONLY).
//
// The reference's default BigInteger is boost::multiprecision::cpp_int
// (reference include/qimcifa.hpp:77-78,94).  Boost is not present in this image, so the
// UNMODIFIED reference sources are compiled against this header instead (oracle/Makefile
// -> oracle/_ref/).  It provides only the operator surface the reference drivers use
// (SURVEY.md section 8b): a 256-bit wrap-around integer with `unsigned __int128` fast
// paths for operands below 2^128 (Boost itself takes a double-limb fast path there), so
// that the timed CPU baseline is not a straw man.  The arithmetic speed is this header's,
// not Boost's; every report that uses oracle/_ref says so.
//
// Optional count (QCF_SHIM_COUNT=<file>, used by bench.py's reference arm): the number of hot-line
// `N % g` evaluations, kept in thread-local counters.
// Optional trace (used to generate tests/golden/ref_traces.json): when the environment
// variable QCF_SHIM_TRACE names a file, the first cpp_int read from a stream is taken as
// "the number to factor" and the right operand of every `N % g` with that left operand
// (the hot line, reference include/qimcifa.hpp:369) is logged; a summary is written at exit.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

namespace boost {
namespace multiprecision {

class cpp_int {
public:
    typedef unsigned __int128 u128;
    uint64_t w[4];

    cpp_int() { w[0] = w[1] = w[2] = w[3] = 0; }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0> cpp_int(T v)
    {
        w[0] = (uint64_t)v;
        const uint64_t ext = (std::is_signed<T>::value && (v < 0)) ? ~0ULL : 0ULL;
        w[1] = w[2] = w[3] = ext;
    }
    static cpp_int from128(u128 v)
    {
        cpp_int r;
        r.w[0] = (uint64_t)v;
        r.w[1] = (uint64_t)(v >> 64);
        return r;
    }

    bool small() const { return !(w[2] | w[3]); }
    u128 lo128() const { return ((u128)w[1] << 64) | w[0]; }
    bool is_zero() const { return !(w[0] | w[1] | w[2] | w[3]); }

    explicit operator bool() const { return !is_zero(); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    explicit operator T() const
    {
        return (T)w[0];
    }
    template <typename T> T convert_to() const
    {
        T r = 0;
        for (int i = 3; i >= 0; --i) {
            r = r * (T)18446744073709551616.0 + (T)w[i];
        }
        return r;
    }

    static int cmp(const cpp_int& a, const cpp_int& b)
    {
        for (int i = 3; i >= 0; --i) {
            if (a.w[i] != b.w[i]) {
                return (a.w[i] < b.w[i]) ? -1 : 1;
            }
        }
        return 0;
    }

    cpp_int& operator+=(const cpp_int& o)
    {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)w[i] + o.w[i];
            w[i] = (uint64_t)c;
            c >>= 64;
        }
        return *this;
    }
    cpp_int& operator-=(const cpp_int& o)
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 4; ++i) {
            const u128 d = (u128)w[i] - o.w[i] - borrow;
            w[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1U;
        }
        return *this;
    }
    cpp_int& operator*=(const cpp_int& o)
    {
        if (!(w[1] | w[2] | w[3] | o.w[1] | o.w[2] | o.w[3])) {
            const u128 p = (u128)w[0] * o.w[0];
            w[0] = (uint64_t)p;
            w[1] = (uint64_t)(p >> 64);
            return *this;
        }
        uint64_t r[4] = { 0, 0, 0, 0 };
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; i + j < 4; ++j) {
                c += (u128)w[i] * o.w[j] + r[i + j];
                r[i + j] = (uint64_t)c;
                c >>= 64;
            }
        }
        for (int i = 0; i < 4; ++i) {
            w[i] = r[i];
        }
        return *this;
    }
    static void divmod(const cpp_int& a, const cpp_int& b, cpp_int& q, cpp_int& r)
    {
        if (a.small() && b.small()) {
            const u128 x = a.lo128(), y = b.lo128();
            if (!(a.w[1] | b.w[1])) {
                q = cpp_int(a.w[0] / b.w[0]);
                r = cpp_int(a.w[0] % b.w[0]);
                return;
            }
            q = from128(x / y);
            r = from128(x % y);
            return;
        }
        // rare wide case: binary long division
        q = cpp_int();
        r = cpp_int();
        for (int i = 255; i >= 0; --i) {
            r <<= 1;
            r.w[0] |= (a.w[i >> 6] >> (i & 63)) & 1U;
            if (cmp(r, b) >= 0) {
                r -= b;
                q.w[i >> 6] |= 1ULL << (i & 63);
            }
        }
    }
    cpp_int& operator/=(const cpp_int& o)
    {
        cpp_int q, r;
        divmod(*this, o, q, r);
        return *this = q;
    }
    cpp_int& operator%=(const cpp_int& o)
    {
        cpp_int q, r;
        divmod(*this, o, q, r);
        return *this = r;
    }
    cpp_int& operator<<=(unsigned s)
    {
        for (; s >= 64; s -= 64) {
            w[3] = w[2];
            w[2] = w[1];
            w[1] = w[0];
            w[0] = 0;
        }
        if (s) {
            w[3] = (w[3] << s) | (w[2] >> (64 - s));
            w[2] = (w[2] << s) | (w[1] >> (64 - s));
            w[1] = (w[1] << s) | (w[0] >> (64 - s));
            w[0] <<= s;
        }
        return *this;
    }
    cpp_int& operator>>=(unsigned s)
    {
        for (; s >= 64; s -= 64) {
            w[0] = w[1];
            w[1] = w[2];
            w[2] = w[3];
            w[3] = 0;
        }
        if (s) {
            w[0] = (w[0] >> s) | (w[1] << (64 - s));
            w[1] = (w[1] >> s) | (w[2] << (64 - s));
            w[2] = (w[2] >> s) | (w[3] << (64 - s));
            w[3] >>= s;
        }
        return *this;
    }
    cpp_int& operator&=(const cpp_int& o)
    {
        for (int i = 0; i < 4; ++i) {
            w[i] &= o.w[i];
        }
        return *this;
    }
    cpp_int& operator++() { return *this += cpp_int(1U); }
    cpp_int& operator--() { return *this -= cpp_int(1U); }
    bool operator!() const { return is_zero(); }
};

// ---- lock-free count of the hot `N % g` (QCF_SHIM_COUNT=<file>) ---------------
// Thread-local counters, added to a global at thread exit, so counting costs one
// increment per guess and no shared cache line; the total is written at process exit.
struct shim_count_t {
    bool enabled = false;
    std::string path;
    std::atomic<uint64_t> total{ 0 };
    shim_count_t()
    {
        const char* p = getenv("QCF_SHIM_COUNT");
        if (p && *p) {
            enabled = true;
            path = p;
        }
    }
    ~shim_count_t()
    {
        if (enabled) {
            FILE* f = fopen(path.c_str(), "w");
            if (f) {
                fprintf(f, "%llu\n", (unsigned long long)total.load());
                fclose(f);
            }
        }
    }
};
inline shim_count_t& shim_count()
{
    static shim_count_t c;
    return c;
}
struct shim_tl_count_t {
    uint64_t n = 0;
    ~shim_tl_count_t() { shim_count().total += n; }
};
inline shim_tl_count_t& shim_tl_count()
{
    static thread_local shim_tl_count_t c;
    return c;
}

// ---- trace of the hot `N % g` ------------------------------------------------
struct shim_trace_t {
    bool enabled = false, have_n = false, counting = false;
    cpp_int n;
    std::string path;
    std::mutex m;
    uint64_t count = 0, hash = 1469598103934665603ULL;
    std::vector<cpp_int> first;
    cpp_int last;
    shim_trace_t()
    {
        const char* p = getenv("QCF_SHIM_TRACE");
        if (p && *p) {
            enabled = true;
            path = p;
        }
    }
    void log(const cpp_int& g)
    {
        std::lock_guard<std::mutex> lock(m);
        ++count;
        for (int i = 0; i < 16; ++i) {
            hash ^= (g.w[i >> 3] >> (8 * (i & 7))) & 0xFFU;
            hash *= 1099511628211ULL;
        }
        if (first.size() < 64) {
            first.push_back(g);
        }
        last = g;
    }
    ~shim_trace_t();
};
inline shim_trace_t& shim_trace()
{
    static shim_trace_t t;
    return t;
}

#define QCF_SHIM_BINOP(op)                                                                                             \
    inline cpp_int operator op(const cpp_int& a, const cpp_int& b)                                                     \
    {                                                                                                                  \
        cpp_int r(a);                                                                                                  \
        r op## = b;                                                                                                    \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline cpp_int operator op(const cpp_int& a, T b)                                                                  \
    {                                                                                                                  \
        cpp_int r(a);                                                                                                  \
        r op## = cpp_int(b);                                                                                           \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline cpp_int operator op(T a, const cpp_int& b)                                                                  \
    {                                                                                                                  \
        cpp_int r(a);                                                                                                  \
        r op## = b;                                                                                                    \
        return r;                                                                                                      \
    }
QCF_SHIM_BINOP(+)
QCF_SHIM_BINOP(-)
QCF_SHIM_BINOP(*)
QCF_SHIM_BINOP(/)
QCF_SHIM_BINOP(&)
#undef QCF_SHIM_BINOP

inline cpp_int operator%(const cpp_int& a, const cpp_int& b)
{
    shim_trace_t& t = shim_trace();
    if (t.enabled && t.have_n && (cpp_int::cmp(a, t.n) == 0)) {
        t.log(b);
    }
    if (t.counting && (cpp_int::cmp(a, t.n) == 0)) {
        ++shim_tl_count().n;
    }
    cpp_int r(a);
    r %= b;
    return r;
}
template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
inline cpp_int operator%(const cpp_int& a, T b)
{
    cpp_int r(a);
    r %= cpp_int(b);
    return r;
}
template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
inline cpp_int operator%(T a, const cpp_int& b)
{
    cpp_int r(a);
    r %= b;
    return r;
}

inline cpp_int operator>>(const cpp_int& a, unsigned s)
{
    cpp_int r(a);
    r >>= s;
    return r;
}
inline cpp_int operator<<(const cpp_int& a, unsigned s)
{
    cpp_int r(a);
    r <<= s;
    return r;
}

#define QCF_SHIM_CMP(op)                                                                                               \
    inline bool operator op(const cpp_int& a, const cpp_int& b) { return cpp_int::cmp(a, b) op 0; }                    \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline bool operator op(const cpp_int& a, T b)                                                                     \
    {                                                                                                                  \
        return cpp_int::cmp(a, cpp_int(b)) op 0;                                                                       \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline bool operator op(T a, const cpp_int& b)                                                                     \
    {                                                                                                                  \
        return cpp_int::cmp(cpp_int(a), b) op 0;                                                                       \
    }
QCF_SHIM_CMP(<)
QCF_SHIM_CMP(<=)
QCF_SHIM_CMP(>)
QCF_SHIM_CMP(>=)
QCF_SHIM_CMP(==)
QCF_SHIM_CMP(!=)
#undef QCF_SHIM_CMP

inline std::string to_string(cpp_int v)
{
    if (v.is_zero()) {
        return "0";
    }
    std::string s;
    const cpp_int ten19(10000000000000000000ULL);
    while (!v.is_zero()) {
        cpp_int q, r;
        cpp_int::divmod(v, ten19, q, r);
        uint64_t d = r.w[0];
        v = q;
        for (int i = 0; i < 19; ++i) {
            s.push_back((char)('0' + (d % 10U)));
            d /= 10U;
            if (v.is_zero() && !d) {
                break;
            }
        }
    }
    return std::string(s.rbegin(), s.rend());
}

inline std::ostream& operator<<(std::ostream& os, const cpp_int& v) { return os << to_string(v); }

inline std::istream& operator>>(std::istream& is, cpp_int& v)
{
    std::string tok;
    is >> tok;
    cpp_int r;
    for (char c : tok) {
        if ((c < '0') || (c > '9')) {
            break;
        }
        r *= cpp_int(10U);
        r += cpp_int((unsigned)(c - '0'));
    }
    v = r;
    shim_trace_t& t = shim_trace();
    if ((t.enabled || shim_count().enabled) && !t.have_n) {
        t.n = v;
        t.have_n = true;
        t.counting = shim_count().enabled;
    }
    return is;
}

inline shim_trace_t::~shim_trace_t()
{
    if (!enabled) {
        return;
    }
    FILE* f = fopen(path.c_str(), "w");
    if (!f) {
        return;
    }
    fprintf(f, "{\"n\": \"%s\", \"count\": %llu, \"fnv1a64\": \"%016llx\", \"last\": \"%s\", \"first\": [", to_string(n).c_str(),
        (unsigned long long)count, (unsigned long long)hash, to_string(last).c_str());
    for (size_t i = 0; i < first.size(); ++i) {
        fprintf(f, "%s\"%s\"", i ? ", " : "", to_string(first[i]).c_str());
    }
    fprintf(f, "]}\n");
    fclose(f);
}

} // namespace multiprecision
} // namespace boost
