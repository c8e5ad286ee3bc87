// tests/hostsim/filter_host.cpp -- TEST INFRASTRUCTURE ONLY: host build of the product's block filters
// (lrzip_next_b200/csrc/filters.cuh), so that the CPU-only container can check them against the reference's own
// converters (oracle/_ref/liblzmaref.so: z7_BranchConv_*_Enc, z7_BranchConvSt_X86_Enc, Delta_Encode).
#include "../../lrzip_next_b200/csrc/filters.cuh"
#include <vector>

using namespace lrz;

// One stream block, in place, the way filters.cu does it (word by word / one serial pass / out[i] = in[i] - in[i - d]).
extern "C" int hostsim_filter_block(int filter, int delta, uint8_t *buf, int64_t n)
{
	if (!flt::supported(filter))
		return -1;
	if (flt::wordwise(filter)) {
		for (int64_t w = 0; w < (n >> 2); w++) {
			uint8_t *p = buf + 4 * w;
			const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
			const uint32_t c = flt::conv_word(filter, v, (uint32_t)(4 * w));
			p[0] = (uint8_t)c;
			p[1] = (uint8_t)(c >> 8);
			p[2] = (uint8_t)(c >> 16);
			p[3] = (uint8_t)(c >> 24);
		}
	} else if (filter == flt::kX86)
		flt::x86_encode(buf, (size_t)n);
	else if (filter == flt::kARMT)
		flt::armt_encode(buf, (size_t)n);
	else if (filter == flt::kRISCV)
		flt::riscv_convert(buf, (size_t)n, true);
	else if (filter == flt::kIA64) {
		for (int64_t o = 0; o + 16 <= n; o += 16)
			flt::ia64_bundle(buf + o, (uint32_t)o);
	}
	else if (filter == flt::kDelta) {
		if (delta < 1 || delta > 256)
			return -1;
		std::vector<uint8_t> in(buf, buf + n);
		for (int64_t i = 0; i < n; i++)
			buf[i] = (uint8_t)(in[i] - (i >= delta ? in[i - delta] : 0));
	}
	return 0;
}

// The inverse converters (decode path: z7_BranchConv_*_Dec, z7_BranchConvSt_X86_Dec, Delta_Decode).
extern "C" int hostsim_unfilter_block(int filter, int delta, uint8_t *buf, int64_t n)
{
	if (!flt::supported(filter))
		return -1;
	if (flt::wordwise(filter)) {
		for (int64_t w = 0; w < (n >> 2); w++) {
			uint8_t *p = buf + 4 * w;
			const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
			const uint32_t c = flt::conv_word(filter, v, (uint32_t)(4 * w), false);
			p[0] = (uint8_t)c;
			p[1] = (uint8_t)(c >> 8);
			p[2] = (uint8_t)(c >> 16);
			p[3] = (uint8_t)(c >> 24);
		}
	} else if (filter == flt::kX86)
		flt::x86_convert(buf, (size_t)n, false);
	else if (filter == flt::kARMT)
		flt::armt_convert(buf, (size_t)n, false);
	else if (filter == flt::kRISCV)
		flt::riscv_convert(buf, (size_t)n, false);
	else if (filter == flt::kIA64) {
		for (int64_t o = 0; o + 16 <= n; o += 16)
			flt::ia64_bundle(buf + o, (uint32_t)o, false);
	} else if (filter == flt::kDelta) {
		if (delta < 1 || delta > 256)
			return -1;
		for (int64_t i = delta; i < n; i++)
			buf[i] = (uint8_t)(buf[i] + buf[i - delta]);
	}
	return 0;
}
