// Host build of the device DEFLATE decoder (vargeno_b200/csrc/vgb_inflate.cuh compiled with one "lane"): test harness only,
// loaded by tests/test_inflate_host.py through ctypes and compared with zlib.  Not part of libvgb200.so.
#include "../../vargeno_b200/csrc/vgb_inflate.cuh"

extern "C" int vgb_host_inflate(const unsigned char *in, unsigned long long in_len, unsigned char *out, unsigned out_cap, unsigned *out_len)
{
	static vgb::InflateWarp t;
	return vgb::inflate_block(in, in_len, out, out_cap, t, out_len);
}
