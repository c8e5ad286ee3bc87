// filters.cuh -- the pre-compression filters of the reference's compthread (src/stream.c:1587-1628),
// applied to every stream-1 block on its own, from position 0 of the block, before the lz4 gate and the backend:
//
//   --delta N   Delta_Encode          src/lzma/C/Delta.c    out[i] = in[i] - in[i - N]   (bytes before the block count as 0)
//   --arm       z7_BranchConv_ARM_Enc   src/lzma/C/Bra.c:127-155   BL (cond = always): 24-bit word offset -> absolute
//   --arm64     z7_BranchConv_ARM64_Enc src/lzma/C/Bra.c:75-124    BL imm26 and ADRP (+-512 MiB reach only)
//   --ppc       z7_BranchConv_PPC_Enc   src/lzma/C/Bra.c:158-195   "bl" (opcode 18, AA = 0, LK = 1), big endian
//   --sparc     z7_BranchConv_SPARC_Enc src/lzma/C/Bra.c:202-257   "call" with a sign-extended 22-bit reach, big endian
//   --x86       z7_BranchConvSt_X86_Enc src/lzma/C/Bra86.c         E8 / E9 rel32, the classic BCJ state machine
//   --armt      z7_BranchConv_ARMT_Enc  src/lzma/C/Bra.c:260-338   Thumb BL pairs (F000 F800), 2-byte units
//   --ia64      z7_BranchConv_IA64_Enc  src/lzma/C/BraIA64.c / Bra.c:343-420   br.call slots of 16-byte bundles
//   --riscv     z7_BranchConv_RISCV_Enc src/lzma/C/Bra.c:423-720   JAL (rd = ra / t0) and AUIPC + I/S-type pairs
//
// The four RISC converters touch aligned 32-bit words independently of each other (a word's new value depends on the
// word and on its offset only), so they are one thread per word; Delta is one thread per byte; the x86 converter
// carries a few bits of state from byte to byte and is run by one thread per block (a stream block is 10 MiB: ~0.1 s,
// beside a block encode of seconds), and so is the Thumb converter (a converted pair hides the half-word after it from
// the scan).  IA64 works on 16-byte bundles independently: one thread per bundle.  RISC-V (JAL, AUIPC pairs) moves
// 4, 6 or 8 bytes ahead depending on what it finds, so it is serial per block like x86.
//
// The converters are stated from the instruction formats; the CPU tests check them byte for byte against the
// reference's own functions (the test-side build of the reference's LZMA SDK exports them), the GPU tests against whole archives of the
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

FLT_FN bool supported(int f) { return f == kNone || f == kX86 || f == kARM || f == kARMT || f == kPPC || f == kSPARC || f == kIA64 || f == kARM64 || f == kRISCV || f == kDelta; }
FLT_FN bool serial(int f) { return f == kX86 || f == kARMT || f == kRISCV; } // the scan's state runs through the block
FLT_FN bool wordwise(int f) { return f == kARM || f == kPPC || f == kSPARC || f == kARM64; }

FLT_FN uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

// ARM: 0xEB in the top byte = BL with condition "always".  Its 24-bit field counts words from the instruction two
// ahead; the filter makes it the absolute word index of the target.
// (every converter below takes `enc`: true adds the position -- compress side -- false subtracts it again -- the
// decode path, z7_BranchConv_*_Dec; the instruction test reads bits the conversion never changes)
FLT_FN uint32_t conv_arm(uint32_t w, uint32_t off, bool enc)
{
	if ((w >> 24) != 0xEBu)
		return w;
	const uint32_t c = (off + 8) >> 2;
	return ((enc ? w + c : w - c) & 0x00ffffffu) | 0xEB000000u;
}

// ARM64: BL (top six bits 100101, imm26 in words from the instruction itself) and ADRP (1 immlo 10000 immhi Rd: a
// 21-bit page delta, converted only while it stays within +-2^17 pages so that data that merely looks like ADRP is
// rarely touched).
FLT_FN uint32_t conv_arm64(uint32_t w, uint32_t off, bool enc)
{
	if ((w & 0xfc000000u) == 0x94000000u)
		return ((enc ? w + (off >> 2) : w - (off >> 2)) & 0x03ffffffu) | 0x94000000u;
	if ((w & 0x9f000000u) != 0x90000000u)
		return w;
	const uint32_t flag = 1u << 20, mask = (1u << 24) - (flag << 1);
	uint32_t v = (w - 0x90000000u) + flag; // immhi biased by 2^17 pages: in range <=> bits 21..23 clear
	if (v & mask)
		return w;
	uint32_t z = (v & 0xffffffe0u) | (v >> 26); // immhi (biased) : immlo, as a number shifted left by 3
	const uint32_t pg = (off >> 9) & ~7u;       // this instruction's page number, same scale
	z = enc ? z + pg : z - pg;
	v = (v & 0x1fu) | 0x90000000u | (z << 26);  // Rd, opcode, new immlo
	v |= 0x00ffffe0u & ((z & ((flag << 1) - 1)) - flag); // new immhi, bias removed
	return v;
}

// PowerPC: "bl target" = opcode 18 with AA = 0, LK = 1; LI (24 bits, in words) is relative to the instruction.
FLT_FN uint32_t conv_ppc(uint32_t w, uint32_t off, bool enc)
{
	uint32_t v = bswap32(w);
	if ((v & 0xfc000003u) != 0x48000001u)
		return w;
	v = ((enc ? v + off : v - off) & 0x03ffffffu) | 0x48000000u;
	return bswap32(v);
}

// SPARC: "call" = 01 + disp30 (words).  Converted when the displacement is a sign-extended 22-bit number, i.e. the
// top ten bits are 01 00000000 or 01 11111111; the absolute target keeps that form (sign re-extended from bit 22).
FLT_FN uint32_t conv_sparc(uint32_t w, uint32_t off, bool enc)
{
	uint32_t v = bswap32(w);
	const uint32_t top = v >> 22;
	if (top != 0x100u && top != 0x1ffu)
		return w;
	uint32_t d = (enc ? (v << 2) + off : (v << 2) - off) >> 2; // absolute (or relative again) word index, 30 bits
	d = (((0u - ((d >> 22) & 1u)) << 22) & 0x3fffffffu) | (d & 0x3fffffu) | 0x40000000u;
	return bswap32(d);
}

FLT_FN uint32_t conv_word(int f, uint32_t w, uint32_t off, bool enc = true)
{
	switch (f) {
	case kARM:
		return conv_arm(w, off, enc);
	case kARM64:
		return conv_arm64(w, off, enc);
	case kPPC:
		return conv_ppc(w, off, enc);
	case kSPARC:
		return conv_sparc(w, off, enc);
	default:
		return w;
	}
}

// x86: E8 (call) / E9 (jmp) followed by a rel32 whose top byte is 00 or FF (a plausible near target) becomes
// absolute, unless one of the three bytes before it was itself such an opcode in a position that makes this one more
// likely to be an operand byte (`mask` remembers that); when the converted value's relevant byte again looks like a
// sign byte the value is folded once more so that the transform stays reversible.  In place, whole block, ip = 0.
FLT_FN bool x86_sign_byte(uint32_t b) { return b == 0 || b == 0xff; }

FLT_FN void x86_convert(uint8_t *buf, size_t n, bool enc)
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
				dest = enc ? src + (uint32_t)(i + 5) : src - (uint32_t)(i + 5);
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

FLT_FN void x86_encode(uint8_t *buf, size_t n) { x86_convert(buf, n, true); }

// ARM Thumb: BL is a pair of half-words 11110 imm11(high) / 11111 imm11(low); the 22-bit field counts half-words from
// the instruction after the pair's first half-word + 2 (i + 4).  After a converted pair the scan moves past it, so the
// second half-word is never taken for the start of another pair: serial, in place, whole block.
FLT_FN void armt_convert(uint8_t *buf, size_t n, bool enc)
{
	for (size_t i = 0; i + 4 <= n; i += 2) {
		if ((buf[i + 1] & 0xf8) != 0xf0 || (buf[i + 3] & 0xf8) != 0xf8)
			continue;
		uint32_t v = (((uint32_t)buf[i + 1] & 7) << 19) | ((uint32_t)buf[i] << 11) | (((uint32_t)buf[i + 3] & 7) << 8) | buf[i + 2];
		v = (enc ? (v << 1) + ((uint32_t)i + 4) : (v << 1) - ((uint32_t)i + 4)) >> 1;
		buf[i + 1] = (uint8_t)(0xf0 | ((v >> 19) & 7));
		buf[i] = (uint8_t)(v >> 11);
		buf[i + 3] = (uint8_t)(0xf8 | ((v >> 8) & 7));
		buf[i + 2] = (uint8_t)v;
		i += 2;
	}
}

FLT_FN void armt_encode(uint8_t *buf, size_t n) { armt_convert(buf, n, true); }

// RISC-V.  Instructions start on 2-byte boundaries; two kinds are converted (everything little endian on the way in):
//   JAL with rd = x1 (ra) or x5 (t0): byte 0 = 0xEF and bits 8, 10, 11 clear.  Its J-type immediate (a byte offset,
//     bits 20..1 scattered over the instruction) becomes the absolute position and is laid out high bits first over
//     the top nibble of byte 1, byte 2, byte 3.
//   AUIPC rd, hi20 followed by a 32-bit instruction that uses rd as rs1 (I/S-type: `addi`, loads, `jalr`, ...), rd
//     other than x0 / x2: the pair's target hi20 + sext(lo12) becomes absolute and goes, big endian, into the second
//     word; the first word becomes an "AUIPC x2" carrying the second instruction's low 20 bits.  A REAL "AUIPC x2"
//     in the input that looks like such a carrier (immediate bits 13:12 set, top five bits not all in {0, 2}) is
//     rotated together with the following word so that the two cases stay apart: decoding is exact.
// The scan moves 2 bytes past a rejected JAL, 4 past a converted JAL or a plain AUIPC x0/x2, 6 past an AUIPC without a
// partner, 8 past a converted pair; the last 6 bytes of the (even-sized) block are never examined.
FLT_FN uint32_t rv_le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
FLT_FN void rv_put_le32(uint8_t *p, uint32_t v)
{
	p[0] = (uint8_t)v;
	p[1] = (uint8_t)(v >> 8);
	p[2] = (uint8_t)(v >> 16);
	p[3] = (uint8_t)(v >> 24);
}
FLT_FN uint32_t rv_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
FLT_FN void rv_put_be32(uint8_t *p, uint32_t v)
{
	p[0] = (uint8_t)(v >> 24);
	p[1] = (uint8_t)(v >> 16);
	p[2] = (uint8_t)(v >> 8);
	p[3] = (uint8_t)v;
}
// second instruction of a pair: a 32-bit encoding (low bits 11) whose rs1 field is `rd`
FLT_FN bool rv_uses_rd(uint32_t second, uint32_t rd) { return (second & 3) == 3 && ((second >> 15) & 31) == rd; }
// "AUIPC x2" that reads as a carrier: immediate bits 13:12 set and a top-five-bits value outside {0, 2}
FLT_FN bool rv_carrier_like(uint32_t w) { return ((w >> 12) & 3) == 3 && ((w >> 27) & 0x1d) != 0; }

FLT_FN void riscv_convert(uint8_t *buf, size_t n, bool enc)
{
	n &= ~(size_t)1;
	if (n <= 6)
		return;
	const size_t lim = n - 6;
	size_t i = 0;
	while (i < lim) {
		uint8_t *p = buf + i;
		const uint32_t op = p[0] & 0x7f;
		if (op != 0x6f && op != 0x17) {
			i += 2;
			continue;
		}
		const uint32_t pc = (uint32_t)i;
		if (op == 0x6f) { // JAL
			if (p[0] != 0xef || (p[1] & 0x0d) != 0) {
				i += 2;
				continue;
			}
			if (enc) {
				const uint32_t w = rv_le32(p);
				uint32_t t = ((w >> 11) & 0x100000u) | ((w >> 20) & 0x7feu) | ((w >> 9) & 0x800u) | (w & 0xff000u);
				t += pc;
				p[1] = (uint8_t)(((t >> 13) & 0xf0) | (p[1] & 0x0f));
				p[2] = (uint8_t)(t >> 9);
				p[3] = (uint8_t)(t >> 1);
			} else {
				uint32_t t = ((uint32_t)(p[1] & 0xf0) << 13) | ((uint32_t)p[2] << 9) | ((uint32_t)p[3] << 1);
				t -= pc;
				const uint32_t w = (uint32_t)p[0] | ((uint32_t)(p[1] & 0x0f) << 8) | ((t << 11) & 0x80000000u) |
						   ((t << 20) & 0x7fe00000u) | ((t << 9) & 0x100000u) | (t & 0xff000u);
				rv_put_le32(p, w);
			}
			i += 4;
			continue;
		}
		// AUIPC
		const uint32_t w = rv_le32(p), rd = (w >> 7) & 31, second = rv_le32(p + 4);
		if (rd != 0 && rd != 2) {
			if (!rv_uses_rd(second, rd)) {
				i += 6;
				continue;
			}
			if (enc) { // a real pair -> carrier + absolute target
				const uint32_t target = (w & 0xfffff000u) + (uint32_t)((int32_t)second >> 20) + pc;
				rv_put_le32(p, (second << 12) | 0x117u);
				rv_put_be32(p + 4, target);
			} else { // the rotated form of a real "AUIPC x2" that looked like a carrier: rotate back
				rv_put_le32(p, (second << 12) | 0x117u);
				rv_put_le32(p + 4, (w & 0xfffff000u) | (second >> 20));
			}
			i += 8;
			continue;
		}
		if (rd == 2 && rv_carrier_like(w)) {
			const uint32_t top = w >> 27;
			if (enc) { // keep it apart from the carriers: rotate it with the word behind it
				rv_put_le32(p, (top << 7) + 0x17u + (second & 0xfffff000u));
				rv_put_le32(p + 4, (w >> 12) | (second << 20));
			} else { // a carrier: rebuild the pair
				const uint32_t target = rv_be32(p + 4) - pc;
				rv_put_le32(p, (top << 7) + 0x17u + ((target + 0x800u) & 0xfffff000u));
				rv_put_le32(p + 4, (w >> 12) | (target << 20));
			}
			i += 8;
			continue;
		}
		i += 4;
	}
}

// IA64: a 16-byte bundle = 5 template bits + three 41-bit slots; the template says which slots hold branch-unit
// instructions.  A slot is converted when it is br.call-like (opcode 5, btype 0): its 21-bit immediate (20 bits at 13,
// sign at 36), in bundles, becomes absolute.  Bundle-local.
FLT_FN void ia64_bundle(uint8_t *b, uint32_t off, bool enc = true)
{
	const uint32_t t = b[0] & 0x1f;
	// slots holding a B unit, by template (10: MIB, 12: MBB, 16: BBB, 18: MMB, 1C: MFB; the odd twin of each ends a group)
	const uint32_t mask = (t == 0x10 || t == 0x11 || t == 0x18 || t == 0x19 || t == 0x1c || t == 0x1d) ? 4u
			      : ((t == 0x12 || t == 0x13) ? 6u : ((t == 0x16 || t == 0x17) ? 7u : 0u));
	for (uint32_t slot = 0, bit = 5; slot < 3; slot++, bit += 41) {
		if (!((mask >> slot) & 1))
			continue;
		const uint32_t bp = bit >> 3, br = bit & 7;
		uint64_t ins = 0;
		for (uint32_t j = 0; j < 6; j++)
			ins |= (uint64_t)b[bp + j] << (8 * j);
		uint64_t x = ins >> br;
		if (((x >> 37) & 0xf) != 5 || ((x >> 9) & 7) != 0)
			continue;
		uint32_t v = (uint32_t)((x >> 13) & 0xfffff) | ((uint32_t)((x >> 36) & 1) << 20);
		v = (enc ? (v << 4) + off : (v << 4) - off) >> 4;
		x &= ~((uint64_t)0x8fffff << 13);
		x |= (uint64_t)(v & 0xfffff) << 13;
		x |= (uint64_t)(v & 0x100000) << (36 - 20);
		ins = (ins & (((uint64_t)1 << br) - 1)) | (x << br);
		for (uint32_t j = 0; j < 6; j++)
			b[bp + j] = (uint8_t)(ins >> (8 * j));
	}
}

} // namespace flt
} // namespace lrz
