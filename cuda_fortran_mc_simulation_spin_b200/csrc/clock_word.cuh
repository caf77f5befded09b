// The per-word (4 sites) update of the q-state clock models with the one-load threshold table, shared by the periodic
// (sixclock.cu: clock_tableall / dual lattice, src/clock/clock_tableall_gpu_m.f90:107-152) and the helical
// (clock_kernels.cuh: clock_gpu_m / clock_gpu_multi_m, src/clock_gpu_m.f90:199-216) pass kernels.
//
// Both passes were bound by the ALU pipe (LOP3 / SHF / ISETP / SEL; profiles/r02b_*): 41-45 thread-instructions per
// site, 22 of them on that pipe.  This form does the per-site work two or four sites per instruction wherever the data
// allow it and moves the rest onto the FMA pipe (IMAD / IMAD.WIDE):
//   * table index  F = n0 + q n1 + q^2 n2 + q^3 n3 + q^4 cur  byte-parallel up to q^3 (< 256 for q <= 6), then in
//     16-bit fields (sites 0, 2 of a word in one register, sites 1, 3 in another);
//   * proposal     k = hi32(V pm + C) by ONE IMAD.WIDE per site, V = the word holding the 16 proposal bits in its high
//     half; its low 32 bits tell whether the 16 bits that are not known yet could change k (tie);
//   * threshold    one LDS.U16 per site from the table T[k][F] (the high 15 bits of the 32-bit threshold);
//   * accept test  (a15 | 0x8000) - th15 in 16-bit fields, two sites per IMAD: bit 15 of a field = reject or tie,
//     field == 0x8000 = tie; the reject bits become a byte mask with one PRMT (sign replication) per word;
//   * ties         the minimum over the fields (DPX VIMNMX3, signed 16 x 2 for the accept test, unsigned 32 for the
//     proposal) is tested once per vector; a flagged word is redone exactly (second Philox block, class table, full
//     33-bit threshold).
//
// RNG contract (v2; CPU restatement: oracle/rng_contract.c, clk_uniform_pair).  One Philox block R serves the four sites
// e = 0..3 of a word, a second block R2 (sub-counter + 4) the low halves:
//   half(w, hs) = hs ? w >> 16 : w & 0xFFFF,   hs = e >> 1
//   accept    a16 = half(R[e & 1], hs),      a16' = half(R2[e & 1], hs)
//   proposal  p16 = half(R[2 + (e & 1)], hs), p16' = half(R2[2 + (e & 1)], hs)
//   U_a = (a16 & 0x7FFF) << 17 | (a16 >> 15) << 16 | a16'         U_p = p16 << 16 | p16'
// (R2 is only evaluated for a tie; the value is the same either way), u = (U + 1) 2^-32 in (0, 1].
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint32_t clk_half(uint32_t w, int hs) { return hs ? w >> 16 : w & 0xFFFFu; }
__device__ __forceinline__ uint32_t clk_word(const uint4& v, int w) { return w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w; }
__device__ __forceinline__ void clk_uniforms(const uint4& R, const uint4& R2, int e, uint32_t& Ua, uint32_t& Up)
{
    const int hs = e >> 1;
    const uint32_t a16 = clk_half(clk_word(R, e & 1), hs), a16b = clk_half(clk_word(R2, e & 1), hs);
    const uint32_t p16 = clk_half(clk_word(R, 2 + (e & 1)), hs), p16b = clk_half(clk_word(R2, 2 + (e & 1)), hs);
    Ua = ((a16 & 0x7FFFu) << 17) | ((a16 >> 15) << 16) | a16b;
    Up = (p16 << 16) | p16b;
}

// {hi, lo} = a * b + c: ONE IMAD.WIDE.U32 when b is an immediate (compile-time q) and c lives in a per-thread register
// pair; with a register multiplier or a uniform / constant addend ptxas splits it into IMAD.WIDE + IADD3 + IADD3.X (seen
// in the SASS), so the kernels are instantiated per q and read the addend from shared memory once (clk_win64)
__device__ __forceinline__ void clk_madwide(uint32_t a, uint32_t b, unsigned long long c, uint32_t& lo, uint32_t& hi)
{
    const unsigned long long r = (unsigned long long)a * b + c;
    lo = (uint32_t)r;
    hi = (uint32_t)(r >> 32);
}
// the kernel stores the window 32 times in shared memory (CLK_WIN_BYTES, in front of its table) and every lane loads its
// own copy: a load from a lane-dependent address is not provably uniform, so the value stays in a per-thread register pair
#define CLK_WIN_BYTES 256
__device__ __forceinline__ void clk_win64_store(void* smem_slots, unsigned long long v)
{
    if (threadIdx.x < 32) reinterpret_cast<unsigned long long*>(smem_slots)[threadIdx.x] = v;
}
__device__ __forceinline__ unsigned long long clk_win64(const void* smem_slots)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"((uint32_t)__cvta_generic_to_shared(smem_slots) + 8u * (threadIdx.x & 31u)));
    return v;
}

// 16-bit fields of the BYTE offset 2 F (< 2 q^5 = 15552) of the table entry for the four sites of a word: n0..n3 = the
// neighbour words in table order, cur = the own word (bytes < q <= 6).  Fe: sites 0 and 2, Fo: sites 1 and 3.
__device__ __forceinline__ void clk_index_fields(uint32_t n0, uint32_t n1, uint32_t n2, uint32_t n3, uint32_t cur, uint32_t q,
                                                 uint32_t& Fe, uint32_t& Fo)
{
    const uint32_t q2 = q * q, q3 = q2 * q;
    const uint32_t A = n0 + q * n1 + q2 * n2;      // bytes < q^3 <= 216
    const uint32_t B = n3 + q * cur;               // bytes < q^2
    Fe = 2u * (A & 0x00FF00FFu) + (2u * q3) * (B & 0x00FF00FFu);
    Fo = 2u * prmt(A, 0u, 0x4341u) + (2u * q3) * prmt(B, 0u, 0x4341u);
}

// One word.  tab: the table T[k][F] (u16) in shared memory, kstride = 2 q^5 bytes; pm = number of proposal cells (q - 1:
// periodic, new = cur + 1 + k mod q; q: helical, new = k); win = pm << 16 as a 64-bit register pair (clk_win64).  Returns
// the new states of the word under the first look; amin / pmin receive the tie indicators of the word: some 16-bit field of
// amin == 0x8000 <=> an accept test is undecided, pmin < 2 win <=> a proposal may depend on the low half (conservative).
template <bool PERIODIC>
__device__ __forceinline__ uint32_t clk_word_fast(uint32_t ow, uint32_t Fe, uint32_t Fo, const uint4& R, const uint8_t* tab, uint32_t kstride,
                                                  uint32_t pm, unsigned long long win, uint32_t q, uint32_t& amin, uint32_t& pmin)
{
    uint32_t k0, k1, k2, k3, l0, l1, l2, l3;
    clk_madwide(R.z << 16, pm, win, l0, k0);   // site 0: low half of R.z
    clk_madwide(R.w << 16, pm, win, l1, k1);   // site 1: low half of R.w
    clk_madwide(R.z, pm, win, l2, k2);         // site 2: high half of R.z (the low half stands in for the unknown bits)
    clk_madwide(R.w, pm, win, l3, k3);         // site 3
    pmin = __vimin3_u32(min(l0, l1), l2, l3);
    const uint32_t t0 = *reinterpret_cast<const uint16_t*>(tab + ((Fe & 0xFFFFu) + k0 * kstride));
    const uint32_t t1 = *reinterpret_cast<const uint16_t*>(tab + ((Fo & 0xFFFFu) + k1 * kstride));
    const uint32_t t2 = *reinterpret_cast<const uint16_t*>(tab + ((Fe >> 16) + k2 * kstride));
    const uint32_t t3 = *reinterpret_cast<const uint16_t*>(tab + ((Fo >> 16) + k3 * kstride));
    const uint32_t De = (R.x | 0x80008000u) - t0 - (t2 << 16);   // fields (a15 | 0x8000) - th15
    const uint32_t Do = (R.y | 0x80008000u) - t1 - (t3 << 16);
    amin = __vmins2(De, Do);
    const uint32_t rej = prmt(De, Do, 0xFBD9u);   // byte e = 0xFF iff site e is rejected (or tied)
    const uint32_t k4 = k0 + (k1 << 8) + (k2 << 16) + (k3 << 24);
    uint32_t nw;
    if (PERIODIC) {
        const uint32_t raw = ow + k4 + 0x01010101u;                         // cur + 1 + k <= 2 q - 2
        const uint32_t ge = ((raw + (0x80u - q) * 0x01010101u) >> 7) & 0x01010101u;   // bytes >= q
        nw = raw - ge * q;
    } else nw = k4;
    return (ow & rej) | (nw & ~rej);
}

// any accept test of the word(s) undecided?  (a field == 0x8000)
__device__ __forceinline__ bool clk_accept_tie(uint32_t amin)
{
    const uint32_t x = amin ^ 0x80008000u;
    return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u;
}
