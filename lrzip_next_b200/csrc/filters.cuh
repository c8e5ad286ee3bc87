// filters.cuh -- the pre-compression filters of the reference's compthread (src/stream.c:1587-1628),
// applied to every stream-1 block on its own, from position 0 of the block, before the lz4 gate and the backend:
//
//   --delta N   Delta_Encode          src/lzma/C/Delta.c    out[i] = in[i] - in[i - N]   (bytes before the block count as 0)
//   --arm       z7_BranchConv_ARM_Enc   src/lzma/C/Bra.c:127-155   BL (cond = always): 24-bit word offset -> absolute
//   --arm64     z7_BranchConv_ARM64_Enc src/lzma/C/Bra.c:75-124    BL imm26 and ADRP (+-512 MiB reach only)
//   --ppc       z7_BranchConv_PPC_Enc   src/lzma/C/Bra.c:158-195   "bl" (opcode 18, AA = 0, LK = 1), big endian
//   --sparc     z7_BranchConv_SPARC_Enc src/lzma/C/Bra.c:202-257   "call" with a sign-extended 22-bit reach, big endian
//   --x86       z7_BranchConvSt_X86_Enc src/lzma/C/Bra86.c         E8 / E9 rel32, the classic BCJ state machine
//
// The four RISC converters touch aligned 32-bit words independently of each other (a word's new value depends on the
// word and on its offset only), so they are one thread per word; Delta is one thread per byte; the x86 converter
// carries a few bits of state from byte to byte and is run by one thread per block (a stream block is 10 MiB: ~0.1 s,
// beside a block encode of seconds).  ARM Thumb, IA64 and RISC-V are not built (LRZGPU_EUNSUPPORTED).
//
// The converters are stated from the instruction formats; the CPU tests check them byte for byte against the
// reference's own functions (oracle/_ref/liblzmaref.so exports them), the GPU tests against whole archives of the
// reference binary.  Compiled for the device (product) and for the host (tests/hostsim only).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FLT_FN __host__ __device__ __forceinline__
#else
#define FLT_FN inline
#endif

namespace lrz {
namespace flt {

// magic byte 16 / control->filter_flag (src/include/lrzip_private.h:389-397)
enum { kNone = 0, kX86 = 1, kARM = 2, kARMT = 3, kPPC = 4, kSPARC = 5, kIA64 = 6, kARM64 = 7, kRISCV = 8, kDelta = 128 };

FLT_FN bool supported(int f) { return f == kNone || f == kX86 || f == kARM || f == kPPC || f == kSPARC || f == kARM64 || f == kDelta; }
FLT_FN bool wordwise(int f) { return f == kARM || f == kPPC || f == kSPARC || f == kARM64; }

FLT_FN uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

// ARM: 0xEB in the top byte = BL with condition "always".  Its 24-bit field counts words from the instruction two
// ahead; the filter makes it the absolute word index of the target.
FLT_FN uint32_t conv_arm(uint32_t w, uint32_t off)
{
	if ((w >> 24) != 0xEBu)
		return w;
	return ((w + ((off + 8) >> 2)) & 0x00ffffffu) | 0xEB000000u;
}

// ARM64: BL (top six bits 100101, imm26 in words from the instruction itself) and ADRP (1 immlo 10000 immhi Rd: a
// 21-bit page delta, converted only while it stays within +-2^17 pages so that data that merely looks like ADRP is
// rarely touched).
FLT_FN uint32_t conv_arm64(uint32_t w, uint32_t off)
{
	if ((w & 0xfc000000u) == 0x94000000u)
		return ((w + (off >> 2)) & 0x03ffffffu) | 0x94000000u;
	if ((w & 0x9f000000u) != 0x90000000u)
		return w;
	const uint32_t flag = 1u << 20, mask = (1u << 24) - (flag << 1);
	uint32_t v = (w - 0x90000000u) + flag; // immhi biased by 2^17 pages: in range <=> bits 21..23 clear
	if (v & mask)
		return w;
	uint32_t z = (v & 0xffffffe0u) | (v >> 26); // immhi (biased) : immlo, as a number shifted left by 3
	z += (off >> 9) & ~7u;                      // + this instruction's page number, same scale
	v = (v & 0x1fu) | 0x90000000u | (z << 26);  // Rd, opcode, new immlo
	v |= 0x00ffffe0u & ((z & ((flag << 1) - 1)) - flag); // new immhi, bias removed
	return v;
}

// PowerPC: "bl target" = opcode 18 with AA = 0, LK = 1; LI (24 bits, in words) is relative to the instruction.
FLT_FN uint32_t conv_ppc(uint32_t w, uint32_t off)
{
	uint32_t v = bswap32(w);
	if ((v & 0xfc000003u) != 0x48000001u)
		return w;
	v = ((v + off) & 0x03ffffffu) | 0x48000000u;
	return bswap32(v);
}

// SPARC: "call" = 01 + disp30 (words).  Converted when the displacement is a sign-extended 22-bit number, i.e. the
// top ten bits are 01 00000000 or 01 11111111; the absolute target keeps that form (sign re-extended from bit 22).
FLT_FN uint32_t conv_sparc(uint32_t w, uint32_t off)
{
	uint32_t v = bswap32(w);
	const uint32_t top = v >> 22;
	if (top != 0x100u && top != 0x1ffu)
		return w;
	uint32_t d = ((v << 2) + off) >> 2;                       // absolute word index, 30 bits
	d = (((0u - ((d >> 22) & 1u)) << 22) & 0x3fffffffu) | (d & 0x3fffffu) | 0x40000000u;
	return bswap32(d);
}

FLT_FN uint32_t conv_word(int f, uint32_t w, uint32_t off)
{
	switch (f) {
	case kARM:
		return conv_arm(w, off);
	case kARM64:
		return conv_arm64(w, off);
	case kPPC:
		return conv_ppc(w, off);
	case kSPARC:
		return conv_sparc(w, off);
	default:
		return w;
	}
}

// x86: E8 (call) / E9 (jmp) followed by a rel32 whose top byte is 00 or FF (a plausible near target) becomes
// absolute, unless one of the three bytes before it was itself such an opcode in a position that makes this one more
// likely to be an operand byte (`mask` remembers that); when the converted value's relevant byte again looks like a
// sign byte the value is folded once more so that the transform stays reversible.  In place, whole block, ip = 0.
FLT_FN bool x86_sign_byte(uint32_t b) { return b == 0 || b == 0xff; }

FLT_FN void x86_encode(uint8_t *buf, size_t n)
{
	if (n < 5)
		return;
	const size_t limit = n - 5;
	uint32_t mask = 0;        // bit k+1: the byte k+1 positions back was E8/E9 (after ageing), bit 4: ... and sign-like
	size_t i = 0, prev = (size_t)0 - 5; // position of the previous opcode candidate
	while (i <= limit) {
		if ((buf[i] & 0xfe) != 0xe8) {
			i++;
			continue;
		}
		const size_t gap = i - prev;
		prev = i;
		if (gap > 5)
			mask = 0;
		else
			for (size_t k = 0; k < gap; k++)
				mask = (mask & 0x77) << 1;
		uint32_t b = buf[i + 4];
		const uint32_t m3 = (mask >> 1) & 7;
		const bool allowed = m3 == 0 || m3 == 1 || m3 == 2 || m3 == 4;
		if (x86_sign_byte(b) && allowed && (mask >> 1) < 0x10) {
			uint32_t src = (b << 24) | ((uint32_t)buf[i + 3] << 16) | ((uint32_t)buf[i + 2] << 8) | buf[i + 1], dest;
			for (;;) {
				dest = src + (uint32_t)(i + 5);
				if (mask == 0)
					break;
				const uint32_t m = mask >> 1, bit = m == 0 ? 0 : (m == 1 ? 1 : (m < 4 ? 2 : 3));
				b = (dest >> (24 - bit * 8)) & 0xff;
				if (!x86_sign_byte(b))
					break;
				src = dest ^ ((1u << (32 - bit * 8)) - 1);
			}
			buf[i + 4] = (uint8_t)(0u - ((dest >> 24) & 1));
			buf[i + 3] = (uint8_t)(dest >> 16);
			buf[i + 2] = (uint8_t)(dest >> 8);
			buf[i + 1] = (uint8_t)dest;
			i += 5;
			mask = 0;
		} else {
			i++;
			mask |= 1;
			if (x86_sign_byte(b))
				mask |= 0x10;
		}
	}
}

} // namespace flt
} // namespace lrz
