// sq_md5.cuh -- RFC 1321 MD5 on the device (Md5.cpp @0x404de0; calcBlockMd5 @0x414d90,
// blockMd5Verify @0x414e00).  One thread digests one buffer: the chain over 64-byte chunks is
// serial, parallelism comes from the 3 x blocks buffers of a wave.
#pragma once
#include "sq_common.cuh"

static __constant__ u32 c_md5k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};

SQ_D u32 rotl(u32 x, int s) { return __funnelshift_l(x, x, s); }
SQ_D void md5_chunk(u32 h[4], const u32 w[16]) {
    u32 a = h[0], b = h[1], c = h[2], d = h[3];
#define MD5_STEP(F, g, s, i) { u32 f = (F) + a + c_md5k[i] + w[g]; a = d; d = c; c = b; b += rotl(f, s); }
#pragma unroll
    for (int i = 0; i < 16; i++) { const int sh[4] = {7, 12, 17, 22}; MD5_STEP((b & c) | (~b & d), i, sh[i & 3], i) }
#pragma unroll
    for (int i = 16; i < 32; i++) { const int sh[4] = {5, 9, 14, 20}; MD5_STEP((d & b) | (~d & c), (5 * i + 1) & 15, sh[i & 3], i) }
#pragma unroll
    for (int i = 32; i < 48; i++) { const int sh[4] = {4, 11, 16, 23}; MD5_STEP(b ^ c ^ d, (3 * i + 5) & 15, sh[i & 3], i) }
#pragma unroll
    for (int i = 48; i < 64; i++) { const int sh[4] = {6, 10, 15, 21}; MD5_STEP(c ^ (b | ~d), (7 * i) & 15, sh[i & 3], i) }
#undef MD5_STEP
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}


// digest of p[0..n) into out[16]
SQ_D void md5_buffer(const u8 *p, u64 n, u8 *out) {
    u32 h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    u32 w[16];
    u64 i = 0;
    if ((((uintptr_t)p) & 15) == 0) {
        const uint4 *p4 = (const uint4 *)p;
        for (; i + 64 <= n; i += 64) {
            uint4 x0 = p4[i / 16], x1 = p4[i / 16 + 1], x2 = p4[i / 16 + 2], x3 = p4[i / 16 + 3];
            w[0] = x0.x; w[1] = x0.y; w[2] = x0.z; w[3] = x0.w; w[4] = x1.x; w[5] = x1.y; w[6] = x1.z; w[7] = x1.w;
            w[8] = x2.x; w[9] = x2.y; w[10] = x2.z; w[11] = x2.w; w[12] = x3.x; w[13] = x3.y; w[14] = x3.z; w[15] = x3.w;
            md5_chunk(h, w);
        }
    } else {
        for (; i + 64 <= n; i += 64) {
            for (int k = 0; k < 16; k++) w[k] = get_u32le(p + i + 4 * k);
            md5_chunk(h, w);
        }
    }
    u8 tail[128];
    u32 rem = (u32)(n - i);
    for (int k = 0; k < 128; k++) tail[k] = 0;
    for (u32 k = 0; k < rem; k++) tail[k] = p[i + k];
    tail[rem] = 0x80;
    u32 tl = rem < 56 ? 64 : 128;
    u64 bits = n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 8 + k] = (u8)(bits >> (8 * k));
    for (u32 o = 0; o < tl; o += 64) {
        for (int k = 0; k < 16; k++) w[k] = get_u32le(tail + o + 4 * k);
        md5_chunk(h, w);
    }
    for (int k = 0; k < 4; k++) put_u32le(out + 4 * k, h[k]);
}
