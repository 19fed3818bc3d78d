// Compat stand-in for <boost/dynamic_bitset.hpp> (TEST INFRASTRUCTURE ONLY).
// Only what the reference uses (reference include/qimcifa.hpp:274-327): size
// constructor, size(), test(), operator[] as an lvalue, operator>>=, copy.
// Blocks are shifted word by word, as Boost does, so the per-step cost the
// reference pays at levels >= 7 is reproduced.
#pragma once

#include <cstddef>
#include <vector>

namespace boost {

template <typename Block = unsigned long> class dynamic_bitset {
    static constexpr size_t BPB = sizeof(Block) * 8;
    std::vector<Block> blocks;
    size_t nbits;

public:
    class reference {
        Block& blk;
        Block mask;

    public:
        reference(Block& b, Block m)
            : blk(b)
            , mask(m)
        {
        }
        reference& operator=(bool v)
        {
            if (v) {
                blk |= mask;
            } else {
                blk &= ~mask;
            }
            return *this;
        }
        operator bool() const { return (blk & mask) != 0; }
    };

    explicit dynamic_bitset(size_t n = 0)
        : blocks((n + BPB - 1) / BPB, 0)
        , nbits(n)
    {
    }
    size_t size() const { return nbits; }
    bool test(size_t i) const { return (blocks[i / BPB] >> (i % BPB)) & 1U; }
    reference operator[](size_t i) { return reference(blocks[i / BPB], (Block)1 << (i % BPB)); }
    bool operator[](size_t i) const { return test(i); }
    dynamic_bitset& operator>>=(size_t s)
    {
        if (!s || blocks.empty()) {
            return *this;
        }
        const size_t div = s / BPB, rem = s % BPB, nb = blocks.size();
        if (div >= nb) {
            for (auto& b : blocks) {
                b = 0;
            }
            return *this;
        }
        for (size_t i = 0; i < nb; ++i) {
            const size_t src = i + div;
            Block v = 0;
            if (src < nb) {
                v = blocks[src] >> rem;
                if (rem && (src + 1 < nb)) {
                    v |= blocks[src + 1] << (BPB - rem);
                }
            }
            blocks[i] = v;
        }
        return *this;
    }
};

} // namespace boost
