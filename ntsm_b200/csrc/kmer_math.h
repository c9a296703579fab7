// kmer_math.h -- the integer arithmetic shared by host and device code.
//
// Reference: vendor/KseqHashIterator.hpp (decode table :114-127, rolling fw/rv :99-100,
// canonical min :102, hash64 :129-139).  The device path works on "stream order" k-mers
// (first base in the LOW bits, as they lie in the packed 2-bit stream); the identities used
// to get back to the reference's values are spelled out at to_reference_orientation().
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NTSM_HD __host__ __device__ __forceinline__
#else
#define NTSM_HD inline
#endif

namespace ntsm {

NTSM_HD uint64_t kmer_mask(unsigned k) { return (1ULL << (2 * k)) - 1; }   // :29

// vendor/KseqHashIterator.hpp:129-139, every stage reduced mod 4^k
NTSM_HD uint64_t hash64(uint64_t x, uint64_t m)
{
	x = (~x + (x << 21)) & m;
	x ^= x >> 24;
	x = (x * 265) & m;      // x + (x<<3) + (x<<8)
	x ^= x >> 14;
	x = (x * 21) & m;       // x + (x<<2) + (x<<4)
	x ^= x >> 28;
	x = (x + (x << 31)) & m;
	return x;
}

// multiplicative inverse of an odd number mod 2^64 (Newton iteration doubles the valid bits)
NTSM_HD uint64_t inv_odd(uint64_t a)
{
	uint64_t x = a;                 // correct to 3 bits
	for (int i = 0; i < 5; ++i) x *= 2 - a * x;
	return x;
}

// inverse of hash64 on [0, 4^k): each stage is an odd multiply, an affine map or a xor-shift
NTSM_HD uint64_t hash64_inv(uint64_t y, uint64_t m)
{
	y = (y * inv_odd((1ULL << 31) + 1)) & m;
	y = y ^ (y >> 28) ^ (y >> 56);
	y = (y * inv_odd(21)) & m;
	y = y ^ (y >> 14) ^ (y >> 28) ^ (y >> 42) ^ (y >> 56);
	y = (y * inv_odd(265)) & m;
	y = y ^ (y >> 24) ^ (y >> 48);
	y = ((y + 1) * inv_odd((1ULL << 21) - 1)) & m;   // stage 1 is x*(2^21-1) - 1
	return y;
}

// reverse the order of the 2-bit groups of a 64-bit word
NTSM_HD uint64_t rev2(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	x = __brevll(x);
#else
	x = ((x >> 32) | (x << 32));
	x = ((x & 0xFFFF0000FFFF0000ULL) >> 16) | ((x & 0x0000FFFF0000FFFFULL) << 16);
	x = ((x & 0xFF00FF00FF00FF00ULL) >> 8) | ((x & 0x00FF00FF00FF00FFULL) << 8);
	x = ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
	x = ((x & 0xCCCCCCCCCCCCCCCCULL) >> 2) | ((x & 0x3333333333333333ULL) << 2);
	x = ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
#endif
	// a full bit reversal also swapped the two bits inside every group: swap them back
	return ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
}

// Stream-order k-mer s (base j of the window in bits [2j,2j+2)) -> the reference's values.
//   fw (first base most significant, :99)  = the 2-bit groups of s in reverse order
//   rv (reverse complement, :100)          = ~s & mask   (complement = 3-c, and reversing
//                                            the reversed order gives stream order back)
NTSM_HD uint64_t stream_to_fw(uint64_t s, unsigned k) { return rev2(s) >> (64 - 2 * k); }
NTSM_HD uint64_t stream_to_rv(uint64_t s, uint64_t m) { return ~s & m; }
NTSM_HD uint64_t fw_to_stream(uint64_t fw, unsigned k) { return rev2(fw << (64 - 2 * k)); }

// cheap mixing of a stream-order k-mer for the pre-filter (NOT the reference hash; the filter
// only has to be free of false negatives, the exact table behind it uses hash64)
NTSM_HD uint32_t filter_mix(uint32_t lo, uint32_t hi)
{
	return (lo + hi * 0x9E3779B1u) * 0x85EBCA6Bu;
}

// decode table, vendor/KseqHashIterator.hpp:114-127
NTSM_HD unsigned nt4(unsigned char c)
{
	switch (c) {
	case 0: case 'A': case 'a': return 0;
	case 1: case 'C': case 'c': return 1;
	case 2: case 'G': case 'g': return 2;
	case 3: case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

}  // namespace ntsm
