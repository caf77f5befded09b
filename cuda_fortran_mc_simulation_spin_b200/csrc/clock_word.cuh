// The per-word (4 sites) update of the q-state clock models with the one-load threshold table, shared by the periodic
// (sixclock.cu: clock_tableall / dual lattice, src/clock/clock_tableall_gpu_m.f90:107-152) and the helical
// (clock_kernels.cuh: clock_gpu_m / clock_gpu_multi_m, src/clock_gpu_m.f90:199-216) pass kernels.
//
// Both passes were bound by the ALU pipe (LOP3 / SHF / ISETP / SEL; profiles/r02b_*): 41-45 thread-instructions per
// site, 22 of them on that pipe.  This form does the per-site work two or four sites per instruction wherever the data
// allow it and moves the rest onto the FMA pipe (IMAD / IMAD.WIDE):
//   * table index  F = n0 + q n1 + q^2 n2 + q^3 n3 + q^4 cur  byte-parallel up to q^3 (< 256 for q <= 6), then in
//     16-bit fields (sites 0, 2 of a word in one register, sites 1, 3 in another);
//   * proposal     the four proposals of a word are the leading base-pm digits of ONE 32-bit uniform (pm = number of
//     proposal cells): {k, W'} = W pm, one IMAD.WIDE per site -- the high word is the digit, the low word the next
//     remainder.  Exact at the full 32 bits: a proposal never needs a second look;
//   * threshold    one LDS.U16 per site from the table T[k][F] (the high 15 bits of the 32-bit threshold);
//   * accept test  (a15 | 0x8000) - th15 in 16-bit fields, two sites per IMAD: bit 15 of a field = reject or tie,
//     field == 0x8000 = tie; the reject bits become a byte mask with one PRMT (sign replication) per word;
//   * ties         the signed 16 x 2 minimum over all fields of the vector (DPX VIMNMX3) is tested once per vector
//     (2^-15 per site with a non-trivial threshold); a vector with a tie is redone exactly (second-stage Philox blocks,
//     class table, full 33-bit threshold).
//
// RNG contract (v3; CPU restatement: oracle/rng_contract.c, clk_vector_uniforms).  A vector = 16 sites = words w = 0..3 of
// sites e = 0..3.  X[0..11] = the words of the Philox blocks with sub-counters 0, 1, 2, Y[0..11] = sub-counters 4, 5, 6.
//   half(x, hs) = hs ? x >> 16 : x & 0xFFFF
//   accept    a16 = half(X[3 w + (e & 1)], e >> 1),  a16' = half(Y[3 w + (e & 1)], e >> 1)
//             U_a = (a16 & 0x7FFF) << 17 | (a16 >> 15) << 16 | a16'        (Y only evaluated for a tie; same value)
//   proposal  W_0 = X[3 w + 2],  W_e = low32(W_(e-1) pm),  U_p(e) = W_e
// u = (U + 1) 2^-32 in (0, 1].  The reference uses the proposal uniform only through ceiling(u (q-1)) (periodic,
// src/clock/clock_tableall_gpu_m.f90:142) or floor(u q) (helical, src/clock_gpu_m.f90:211), i.e. through the digit
// floor(W_e pm / 2^32): the digits of one uniform are independent and uniform up to pm^4 2^-32 = 1.5e-7 (3e-7 for pm = 6).
#pragma once
#include "common.cuh"

__device__ __forceinline__ uint32_t clk_half(uint32_t w, int hs) { return hs ? w >> 16 : w & 0xFFFFu; }
__device__ __forceinline__ uint32_t clk_accept32(uint32_t x, uint32_t y, int hs)
{
    const uint32_t a16 = clk_half(x, hs), a16b = clk_half(y, hs);
    return ((a16 & 0x7FFFu) << 17) | ((a16 >> 15) << 16) | a16b;
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
// periodic, new = cur + 1 + k mod q; q: helical, new = k); xe / xo / xp = X[3 w], X[3 w + 1], X[3 w + 2].  Returns the new
// states of the word under the first look; amin = min(amin, fields): a field == 0x8000 <=> an accept test is undecided.
template <bool PERIODIC>
__device__ __forceinline__ uint32_t clk_word_fast(uint32_t ow, uint32_t Fe, uint32_t Fo, uint32_t xe, uint32_t xo, uint32_t xp, const uint8_t* tab,
                                                  uint32_t kstride, uint32_t pm, uint32_t q, uint32_t& amin)
{
    uint32_t k0, k1, k2, k3, w = xp;
    mulwide(w, pm, w, k0);
    mulwide(w, pm, w, k1);
    mulwide(w, pm, w, k2);
    mulwide(w, pm, w, k3);
    const uint32_t t0 = *reinterpret_cast<const uint16_t*>(tab + ((Fe & 0xFFFFu) + k0 * kstride));
    const uint32_t t1 = *reinterpret_cast<const uint16_t*>(tab + ((Fo & 0xFFFFu) + k1 * kstride));
    const uint32_t t2 = *reinterpret_cast<const uint16_t*>(tab + ((Fe >> 16) + k2 * kstride));
    const uint32_t t3 = *reinterpret_cast<const uint16_t*>(tab + ((Fo >> 16) + k3 * kstride));
    const uint32_t De = (xe | 0x80008000u) - t0 - (t2 << 16);   // fields (a15 | 0x8000) - th15
    const uint32_t Do = (xo | 0x80008000u) - t1 - (t3 << 16);
    amin = __vimin3_s16x2(amin, De, Do);
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

// any accept test undecided?  (a field == 0x8000)
__device__ __forceinline__ bool clk_accept_tie(uint32_t amin)
{
    const uint32_t x = amin ^ 0x80008000u;
    return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u;
}
