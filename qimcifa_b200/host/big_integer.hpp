// Host-side fixed-width BigInteger for the qimcifa drivers.
//
// Keeps the type surface the reference drivers rely on: the reference's default
// BigInteger is boost::multiprecision::cpp_int (reference include/qimcifa.hpp:94) and its own
// fixed-width type stores BIG_INT_BITS/64 little-endian 64-bit words (reference
// include/big_integer.hpp:60-61).  This class is written from scratch: 256 bits (so mid*mid in sqrt()
// fits for 128-bit N), little-endian 64-bit words, trivially copyable, with the operators the drivers
// use -- including the ones the reference's own type lacks (>=, bool, by-reference *= and /=).
// It never runs on the hot path: the GPU kernels use 32-bit limb arithmetic (csrc/qcf_device.cuh).
#pragma once

#include <cstdint>
#include <iostream>
#include <string>
#include <type_traits>

namespace Qimcifa {

class BigInteger {
public:
    static constexpr int WORDS = 4;
    uint64_t word[WORDS];

    BigInteger() { clear(); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0> BigInteger(T v)
    {
        const uint64_t fill = (std::is_signed<T>::value && (v < 0)) ? ~0ULL : 0ULL;
        word[0] = (uint64_t)v;
        for (int i = 1; i < WORDS; ++i) {
            word[i] = fill;
        }
    }

    void clear()
    {
        for (int i = 0; i < WORDS; ++i) {
            word[i] = 0;
        }
    }
    bool isZero() const
    {
        uint64_t acc = 0;
        for (int i = 0; i < WORDS; ++i) {
            acc |= word[i];
        }
        return acc == 0;
    }
    int bitLength() const
    {
        for (int i = WORDS - 1; i >= 0; --i) {
            if (word[i]) {
                return 64 * i + 64 - __builtin_clzll(word[i]);
            }
        }
        return 0;
    }
    bool fitsWords(int n) const
    {
        for (int i = n; i < WORDS; ++i) {
            if (word[i]) {
                return false;
            }
        }
        return true;
    }
    // 32-bit little-endian limbs for the C ABI (include/qimcifa_cuda.h)
    void toLimbs(uint32_t* le, int n) const
    {
        for (int i = 0; i < n; ++i) {
            le[i] = (uint32_t)(word[i >> 1] >> (32 * (i & 1)));
        }
    }
    static BigInteger fromLimbs(const uint32_t* le, int n)
    {
        BigInteger r;
        for (int i = 0; i < n; ++i) {
            r.word[i >> 1] |= (uint64_t)le[i] << (32 * (i & 1));
        }
        return r;
    }

    explicit operator bool() const { return !isZero(); }
    template <typename T, typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    explicit operator T() const
    {
        return (T)word[0];
    }
    template <typename T> T convert_to() const // cpp_int spelling, used by the tuner (reference src/qimcifa_tuner.cpp:143)
    {
        T r = 0;
        for (int i = WORDS - 1; i >= 0; --i) {
            r = r * (T)18446744073709551616.0 + (T)word[i];
        }
        return r;
    }

    static int compare(const BigInteger& a, const BigInteger& b)
    {
        for (int i = WORDS - 1; i >= 0; --i) {
            if (a.word[i] != b.word[i]) {
                return (a.word[i] > b.word[i]) ? 1 : -1;
            }
        }
        return 0;
    }

    BigInteger& operator+=(const BigInteger& o)
    {
        unsigned __int128 carry = 0;
        for (int i = 0; i < WORDS; ++i) {
            carry += (unsigned __int128)word[i] + o.word[i];
            word[i] = (uint64_t)carry;
            carry >>= 64;
        }
        return *this;
    }
    BigInteger& operator-=(const BigInteger& o)
    {
        uint64_t borrow = 0;
        for (int i = 0; i < WORDS; ++i) {
            const uint64_t a = word[i], b = o.word[i];
            const uint64_t d = a - b - borrow;
            borrow = (a < b) || ((a == b) && borrow) ? 1 : 0;
            word[i] = d;
        }
        return *this;
    }
    BigInteger& operator*=(const BigInteger& o)
    {
        uint64_t acc[WORDS] = { 0, 0, 0, 0 };
        for (int i = 0; i < WORDS; ++i) {
            if (!word[i]) {
                continue;
            }
            unsigned __int128 carry = 0;
            for (int j = 0; i + j < WORDS; ++j) {
                carry += (unsigned __int128)word[i] * o.word[j] + acc[i + j];
                acc[i + j] = (uint64_t)carry;
                carry >>= 64;
            }
        }
        for (int i = 0; i < WORDS; ++i) {
            word[i] = acc[i];
        }
        return *this;
    }
    BigInteger& operator<<=(unsigned s)
    {
        const unsigned ws = s / 64, bs = s % 64;
        for (int i = WORDS - 1; i >= 0; --i) {
            uint64_t v = 0;
            if (i >= (int)ws) {
                v = word[i - ws] << bs;
                if (bs && (i - (int)ws - 1 >= 0)) {
                    v |= word[i - ws - 1] >> (64 - bs);
                }
            }
            word[i] = v;
        }
        return *this;
    }
    BigInteger& operator>>=(unsigned s)
    {
        const unsigned ws = s / 64, bs = s % 64;
        for (int i = 0; i < WORDS; ++i) {
            uint64_t v = 0;
            if (i + ws < (unsigned)WORDS) {
                v = word[i + ws] >> bs;
                if (bs && (i + ws + 1 < (unsigned)WORDS)) {
                    v |= word[i + ws + 1] << (64 - bs);
                }
            }
            word[i] = v;
        }
        return *this;
    }
    BigInteger& operator&=(const BigInteger& o)
    {
        for (int i = 0; i < WORDS; ++i) {
            word[i] &= o.word[i];
        }
        return *this;
    }

    // quotient and remainder; schoolbook bit-serial above 128 bits, hardware division below
    static void divMod(const BigInteger& num, const BigInteger& den, BigInteger& quo, BigInteger& rem)
    {
        if (num.fitsWords(2) && den.fitsWords(2) && !den.isZero()) {
            const unsigned __int128 a = ((unsigned __int128)num.word[1] << 64) | num.word[0];
            const unsigned __int128 b = ((unsigned __int128)den.word[1] << 64) | den.word[0];
            const unsigned __int128 q = a / b, r = a % b;
            quo.clear();
            rem.clear();
            quo.word[0] = (uint64_t)q;
            quo.word[1] = (uint64_t)(q >> 64);
            rem.word[0] = (uint64_t)r;
            rem.word[1] = (uint64_t)(r >> 64);
            return;
        }
        BigInteger q, r;
        for (int bit = num.bitLength() - 1; bit >= 0; --bit) {
            r <<= 1U;
            r.word[0] |= (num.word[bit >> 6] >> (bit & 63)) & 1ULL;
            if (compare(r, den) >= 0) {
                r -= den;
                q.word[bit >> 6] |= 1ULL << (bit & 63);
            }
        }
        quo = q;
        rem = r;
    }
    BigInteger& operator/=(const BigInteger& o)
    {
        BigInteger q, r;
        divMod(*this, o, q, r);
        return *this = q;
    }
    BigInteger& operator%=(const BigInteger& o)
    {
        BigInteger q, r;
        divMod(*this, o, q, r);
        return *this = r;
    }
    BigInteger& operator++() { return *this += BigInteger(1U); }
    BigInteger& operator--() { return *this -= BigInteger(1U); }
    BigInteger operator++(int)
    {
        const BigInteger old(*this);
        ++*this;
        return old;
    }
    bool operator!() const { return isZero(); }

    std::string str() const
    {
        if (isZero()) {
            return "0";
        }
        std::string digits;
        BigInteger v(*this);
        const BigInteger chunk(1000000000000000000ULL); // 10^18
        while (!v.isZero()) {
            BigInteger q, r;
            divMod(v, chunk, q, r);
            uint64_t d = r.word[0];
            v = q;
            for (int i = 0; (i < 18) && (d || !v.isZero()); ++i) {
                digits.push_back((char)('0' + d % 10U));
                d /= 10U;
            }
        }
        return std::string(digits.rbegin(), digits.rend());
    }
    static BigInteger parse(const std::string& s)
    {
        BigInteger r;
        for (char c : s) {
            if ((c < '0') || (c > '9')) {
                break;
            }
            r *= BigInteger(10U);
            r += BigInteger((unsigned)(c - '0'));
        }
        return r;
    }
};

#define QIMCIFA_BI_ARITH(OP)                                                                                           \
    inline BigInteger operator OP(const BigInteger& a, const BigInteger& b)                                            \
    {                                                                                                                  \
        BigInteger r(a);                                                                                               \
        r OP## = b;                                                                                                    \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline BigInteger operator OP(const BigInteger& a, T b)                                                            \
    {                                                                                                                  \
        return a OP BigInteger(b);                                                                                     \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline BigInteger operator OP(T a, const BigInteger& b)                                                            \
    {                                                                                                                  \
        return BigInteger(a) OP b;                                                                                     \
    }
QIMCIFA_BI_ARITH(+)
QIMCIFA_BI_ARITH(-)
QIMCIFA_BI_ARITH(*)
QIMCIFA_BI_ARITH(/)
QIMCIFA_BI_ARITH(%)
QIMCIFA_BI_ARITH(&)
#undef QIMCIFA_BI_ARITH

template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
inline BigInteger& operator+=(BigInteger& a, T b)
{
    return a += BigInteger(b);
}
template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
inline BigInteger& operator*=(BigInteger& a, T b)
{
    return a *= BigInteger(b);
}

inline BigInteger operator>>(const BigInteger& a, unsigned s)
{
    BigInteger r(a);
    r >>= s;
    return r;
}
inline BigInteger operator<<(const BigInteger& a, unsigned s)
{
    BigInteger r(a);
    r <<= s;
    return r;
}

#define QIMCIFA_BI_CMP(OP)                                                                                             \
    inline bool operator OP(const BigInteger& a, const BigInteger& b) { return BigInteger::compare(a, b) OP 0; }       \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline bool operator OP(const BigInteger& a, T b)                                                                  \
    {                                                                                                                  \
        return BigInteger::compare(a, BigInteger(b)) OP 0;                                                             \
    }                                                                                                                  \
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>                          \
    inline bool operator OP(T a, const BigInteger& b)                                                                  \
    {                                                                                                                  \
        return BigInteger::compare(BigInteger(a), b) OP 0;                                                             \
    }
QIMCIFA_BI_CMP(<)
QIMCIFA_BI_CMP(<=)
QIMCIFA_BI_CMP(>)
QIMCIFA_BI_CMP(>=)
QIMCIFA_BI_CMP(==)
QIMCIFA_BI_CMP(!=)
#undef QIMCIFA_BI_CMP

inline std::ostream& operator<<(std::ostream& os, const BigInteger& v) { return os << v.str(); }
inline std::istream& operator>>(std::istream& is, BigInteger& v)
{
    std::string tok;
    is >> tok;
    v = BigInteger::parse(tok);
    return is;
}

} // namespace Qimcifa
