// sz3_b200/csrc/zhuf_kernels.cu -- the GPU lossless stage: zstd frames of Huffman-only literal blocks (zhuf.cuh).
//
//   k_zhuf_build    one CTA per 128 KiB block: byte histogram (shared-memory atomics), rank sort of the 256 counters,
//                   Huffman code + zstd tree description (one thread; a few thousand serial steps), bytes of the 4 streams
//   k_zhuf_scan     one CTA: coded / raw decision per block, output offsets (frame and block headers included), total
//   k_zhuf_encode   one CTA per stream: per-thread chunk bit counts, suffix scan, codes OR-ed into a shared-memory bit
//                   buffer (last symbol first, LSB-first), closing bit, byte copy to the stream's place; stream 0's CTA
//                   also writes the block's headers
//
// HBM-bound byte work: the source is read three times (histogram, sizes, encode; the second and third hit L2 for the
// block sizes involved) and the output written once.
#include <cuda_runtime.h>

#include "launch.hpp"
#include "zhuf.cuh"
#include "zhuf_dec.cuh"

namespace sz3b {

constexpr int kZhufThreads = 256;

__global__ void __launch_bounds__(kZhufThreads) k_zhuf_build(const uint8_t *__restrict__ src, uint64_t len,
                                                             ZhufBlockInfo *__restrict__ infos, uint64_t g0) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t hist4[4][256];   // per stream (the block's four segments): their sizes follow from the code
    __shared__ uint8_t ss[256];
    __shared__ uint32_t sf[256];
    __shared__ ZhufScratch scratch;
    __shared__ ZhufBlockInfo info;
    __shared__ unsigned long long bits[4];
    const uint64_t g = g0 + blockIdx.x;
    const uint32_t bl = zhuf_block_len(len, g);
    const uint8_t *p = src + g * kZhufBlock;
    const int tid = threadIdx.x;
    for (int s = 0; s < 4; s++) hist4[s][tid] = 0;
    if (tid < 4) bits[tid] = 0;
    __syncthreads();
    const uint32_t nvec = bl / 16;   // blocks start 16-byte aligned (the stream buffer is, and 128 KiB divides evenly)
    const uint32_t seg = (bl + 3) / 4;
    const uint4 *pv = reinterpret_cast<const uint4 *>(p);
    for (uint32_t i = tid; i < nvec; i += kZhufThreads) {
        const uint4 v = pv[i];
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
        const uint32_t at = i * 16;
        uint32_t s = at / seg;   // a 16-byte group lies in one stream whenever seg is a multiple of 16 (full blocks)
        if (s > 3) s = 3;
        if ((at + 15) / seg == at / seg || s == 3) {
            uint32_t *h = hist4[s];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                atomicAdd(&h[wv[k] & 0xff], 1u);
                atomicAdd(&h[(wv[k] >> 8) & 0xff], 1u);
                atomicAdd(&h[(wv[k] >> 16) & 0xff], 1u);
                atomicAdd(&h[wv[k] >> 24], 1u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t sj = (at + 4 * k + b) / seg;
                    atomicAdd(&hist4[sj > 3 ? 3 : sj][(wv[k] >> (8 * b)) & 0xff], 1u);
                }
            }
        }
    }
    for (uint32_t i = nvec * 16 + tid; i < bl; i += kZhufThreads) {
        const uint32_t sj = i / seg;
        atomicAdd(&hist4[sj > 3 ? 3 : sj][p[i]], 1u);
    }
    __syncthreads();
    hist[tid] = hist4[0][tid] + hist4[1][tid] + hist4[2][tid] + hist4[3][tid];
    __syncthreads();
    // rank sort by (count, symbol): thread s places symbol s
    const uint32_t mine = hist[tid];
    int rank = 0;
    for (int j = 0; j < 256; j++) {
        const uint32_t h = hist[j];
        rank += (h != 0 && (h < mine || (h == mine && j < tid))) ? 1 : 0;
    }
    if (mine) {
        ss[rank] = static_cast<uint8_t>(tid);
        sf[rank] = mine;
    }
    const int n = __syncthreads_count(mine != 0);
    zhuf_build_table(ss, sf, n, scratch, info, tid, kZhufThreads);
    __syncthreads();
    // bits of the four streams under this code: per-stream symbol counts times code lengths (no second pass over the
    // source)
    unsigned long long acc[4];
    {
        const unsigned long long l = info.sym[tid] >> 16;
#pragma unroll
        for (int s = 0; s < 4; s++) acc[s] = l * hist4[s][tid];
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        unsigned long long v = acc[s];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&bits[s], v);
    }
    __syncthreads();
    if (tid < 4) info.sb[tid] = static_cast<uint32_t>(bits[tid] / 8 + 1);
    __syncthreads();
    uint32_t *dst = reinterpret_cast<uint32_t *>(&infos[g]);
    const uint32_t *sp = reinterpret_cast<const uint32_t *>(&info);
    for (uint32_t i = tid; i < sizeof(ZhufBlockInfo) / 4; i += kZhufThreads) dst[i] = sp[i];
}

// Blocks [g0, g1): coded / raw decision and output offsets, continuing from *total (the bytes of the blocks before g0);
// *total is advanced and a copy is left in total_log[0] for the host (which learns the byte range of this slice from it).
__global__ void __launch_bounds__(1024) k_zhuf_scan(ZhufBlockInfo *__restrict__ infos, uint64_t len, uint64_t g0, uint64_t g1,
                                                    unsigned long long *__restrict__ total,
                                                    unsigned long long *__restrict__ total_log) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = *total;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t start = g0; start < g1; start += 1024) {
        const uint64_t g = start + threadIdx.x;
        unsigned long long v = 0, pre = 0;
        bool coded = false;
        if (g < g1) {
            const uint32_t payload = zhuf_block_payload(zhuf_block_len(len, g), infos[g], &coded);
            pre = g % kZhufBlocksPerFrame == 0 ? kZhufFrameHeader : 0;
            v = pre + 3 + payload;
        }
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = warp_sum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long before = carry + (wid ? warp_sum[wid - 1] : 0ull) + x - v;
        if (g < g1) {
            infos[g].coded = coded ? 1u : 0u;
            infos[g].off = before + pre;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = carry_s;
        *total_log = carry_s;   // may be mapped host memory: the host reads it after the event that follows the slice
        __threadfence_system();
    }
}

constexpr uint32_t kZhufBufWords = (kZhufBlock / 4 * kZhufMaxBits + 31) / 32 + 2;   // bits of one stream + closing bit

__global__ void __launch_bounds__(kZhufThreads) k_zhuf_encode(const uint8_t *__restrict__ src, uint64_t len,
                                                              const ZhufBlockInfo *__restrict__ infos,
                                                              uint8_t *__restrict__ out, uint64_t g0) {
    __shared__ uint32_t sym_s[256];
    __shared__ uint32_t buf[kZhufBufWords];
    __shared__ uint32_t warp_sum[kZhufThreads / 32];
    const uint64_t g = g0 + (blockIdx.x >> 2);
    const int s = blockIdx.x & 3;
    const int tid = threadIdx.x;
    const ZhufBlockInfo &bi = infos[g];
    if (s == 0 && tid == 0) zhuf_write_headers(out, len, g, bi);
    uint64_t a, b;
    zhuf_stream_range(len, g, s, &a, &b);
    const uint32_t n = static_cast<uint32_t>(b - a);
    if (!bi.coded) {
        uint8_t *dst = out + bi.off + 3 + (a - g * kZhufBlock);
        for (uint32_t i = tid; i < n; i += kZhufThreads) dst[i] = src[a + i];
        return;
    }
    const uint32_t sb = bi.sb[s];
    uint64_t p = bi.off + 3 + 5 + bi.desc_len + 6;
    for (int k = 0; k < s; k++) p += bi.sb[k];
    sym_s[tid] = bi.sym[tid];
    for (uint32_t i = tid; i < (sb + 3) / 4 + 1 && i < kZhufBufWords; i += kZhufThreads) buf[i] = 0;
    __syncthreads();
    // chunk of this thread, bit count.  Full streams (32768 symbols, 128 per thread, 16-byte aligned) keep their bytes
    // in registers between the two passes.
    const uint32_t per = (n + kZhufThreads - 1) / kZhufThreads;
    const uint32_t c0 = tid * per < n ? tid * per : n, c1 = c0 + per < n ? c0 + per : n;
    const uint8_t *sp = src + a;
    const bool vec = per == 128 && n == 128 * kZhufThreads && ((a & 15) == 0);
    uint4 v[8];
    uint32_t mybits = 0;
    if (vec) {
        const uint4 *pv = reinterpret_cast<const uint4 *>(sp + c0);
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = pv[k];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t wv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int j = 0; j < 4; j++)
                mybits += (sym_s[wv[j] & 0xff] >> 16) + (sym_s[(wv[j] >> 8) & 0xff] >> 16) + (sym_s[(wv[j] >> 16) & 0xff] >> 16) +
                          (sym_s[wv[j] >> 24] >> 16);
        }
    } else {
        for (uint32_t i = c0; i < c1; i++) mybits += sym_s[sp[i]] >> 16;
    }
    // bits written before this chunk = bits of all LATER chunks (the last symbol goes first)
    uint32_t x = mybits;
    const int lane = tid & 31, wid = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kZhufThreads / 32; w++) {
        if (w < wid) before += warp_sum[w];
        total += warp_sum[w];
    }
    // Codes are gathered in a 64-bit register and leave as whole words.  The words strictly inside the chunk's bit
    // range belong to this thread alone (plain stores); the first and the last one are shared with the neighbours.
    const uint32_t o0 = total - (before + x);
    unsigned long long acc = 0;
    uint32_t nacc = o0 & 31, w = o0 >> 5;
    bool first = true;
    auto emit = [&](uint32_t byte) {
        const uint32_t e = sym_s[byte];
        acc |= static_cast<unsigned long long>(e & 0xffffu) << nacc;
        nacc += e >> 16;
        if (nacc >= 32) {
            if (first)
                atomicOr(&buf[w], static_cast<uint32_t>(acc));
            else
                buf[w] = static_cast<uint32_t>(acc);
            first = false;
            acc >>= 32;
            nacc -= 32;
            w++;
        }
    };
    if (vec) {
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            const uint32_t wv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int j = 3; j >= 0; j--) {
                emit(wv[j] >> 24);
                emit((wv[j] >> 16) & 0xff);
                emit((wv[j] >> 8) & 0xff);
                emit(wv[j] & 0xff);
            }
        }
    } else {
        for (uint32_t i = c1; i-- > c0;) emit(sp[i]);
    }
    if (nacc && acc) atomicOr(&buf[w], static_cast<uint32_t>(acc));
    if (tid == 0) atomicOr(&buf[total >> 5], 1u << (total & 31));   // closing bit
    __syncthreads();
    // copy out: bytes up to the first 4-byte boundary of the destination, aligned words, tail bytes
    const uint8_t *bb = reinterpret_cast<const uint8_t *>(buf);
    uint8_t *dst = out + p;
    const uint32_t head = static_cast<uint32_t>((4 - (p & 3)) & 3) < sb ? static_cast<uint32_t>((4 - (p & 3)) & 3) : sb;
    if (tid < static_cast<int>(head)) dst[tid] = bb[tid];
    const uint32_t nwords = (sb - head) / 4;
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
    for (uint32_t i = tid; i < nwords; i += kZhufThreads)
        dw[i] = head ? __funnelshift_r(buf[i], buf[i + 1], 8 * head) : buf[i];
    for (uint32_t i = head + nwords * 4 + tid; i < sb; i += kZhufThreads) dst[i] = bb[i];
}

// Decoder of zhuf-shaped frames (zhuf_dec.cuh): one CTA per block.  Thread 0 reads the tree description (the FSE chain
// over the weights is serial), all threads build the block's decode table in shared memory, then lane 0 of each of
// the four warps decodes one of the four streams -- 32768 dependent table lookups each, so the kernel's time is one
// stream's latency whatever the number of blocks.  Raw blocks are copied.  *bad is set when a block does not decode
// (the caller then hands the whole payload to libzstd).
constexpr int kZhufDecThreads = 128;
constexpr uint32_t kZhufDecStage = 192;
__global__ void __launch_bounds__(kZhufDecThreads) k_zhuf_decode(const uint8_t *__restrict__ cmp, const ZhufDecBlock *__restrict__ blocks,
                                                                 uint8_t *__restrict__ raw, unsigned *__restrict__ bad) {
    __shared__ ZhufDecScratch S;
    __shared__ uint16_t tab[1 << kZhufMaxBits];
    __shared__ uint32_t sb[4], sn[4];
    __shared__ uint8_t sdesc[kZhufDecStage];
    const ZhufDecBlock b = blocks[blockIdx.x];
    const int tid = threadIdx.x;
    if (!b.coded) {
        const uint8_t *s = cmp + b.src;
        uint8_t *d = raw + b.dst;
        for (uint32_t i = tid; i < b.regen; i += kZhufDecThreads) d[i] = s[i];
        return;
    }
    const uint8_t *d = cmp + b.src;
    // the tree description (at most 129 bytes) and the jump table behind it, staged in shared memory: thread 0 reads
    // them bit by bit
    const uint32_t nstage = b.lit < kZhufDecStage ? b.lit : kZhufDecStage;
    for (uint32_t i = tid; i < nstage; i += kZhufDecThreads) sdesc[i] = d[i];
    __syncthreads();
    if (tid == 0) {
        S.ok = zhuf_read_weights(sdesc, nstage, S) ? 1 : 0;
        if (S.ok && (S.desc_len + 6 > nstage || !zhuf_stream_sizes(sdesc, b.lit, S.desc_len, b.regen, sb, sn))) S.ok = 0;
        if (!S.ok) atomicOr(bad, 1u);
    }
    __syncthreads();
    if (!S.ok) return;
    zhuf_dec_table(S, tab, tid, kZhufDecThreads);
    if ((tid & 31) == 0) {
        const int s = tid >> 5;
        const uint8_t *sp = d + S.desc_len + 6;
        uint8_t *dp = raw + b.dst;
        for (int k = 0; k < s; k++) {
            sp += sb[k];
            dp += sn[k];
        }
        if (!zhuf_dec_stream(sp, sb[s], tab, S.maxbits, dp, sn[s])) atomicOr(bad, 2u);
    }
}

void launch_zhuf_decode(const uint8_t *cmp, const ZhufDecBlock *blocks, size_t nblocks, uint8_t *raw, unsigned *bad, cudaStream_t st) {
    if (nblocks) k_zhuf_decode<<<static_cast<unsigned>(nblocks), kZhufDecThreads, 0, st>>>(cmp, blocks, raw, bad);
}

// Tables and stream sizes of every block of src[0, len) (device).  One launch for the whole stream: the kernel's time
// is the latency of one block's table construction, whatever the number of blocks.
void launch_zhuf_build(const uint8_t *src, uint64_t len, ZhufBlockInfo *infos, cudaStream_t st) {
    const uint64_t nblocks = zhuf_num_blocks(len);
    if (nblocks) k_zhuf_build<<<static_cast<unsigned>(nblocks), kZhufThreads, 0, st>>>(src, len, infos, 0);
}

// Offsets and frames of the blocks [g0, g1) into out (device, zhuf_bound(len) bytes).  *total (device) must hold the
// bytes produced for the blocks before g0 (0 for the first slice) and receives the new running total, which is also
// left in *total_log.  Slices must start on frame boundaries (multiples of kZhufBlocksPerFrame).
void launch_zhuf_emit(const uint8_t *src, uint64_t len, uint64_t g0, uint64_t g1, ZhufBlockInfo *infos, uint8_t *out,
                      unsigned long long *total, unsigned long long *total_log, cudaStream_t st) {
    if (g1 <= g0) return;
    k_zhuf_scan<<<1, 1024, 0, st>>>(infos, len, g0, g1, total, total_log);
    k_zhuf_encode<<<static_cast<unsigned>((g1 - g0) * 4), kZhufThreads, 0, st>>>(src, len, infos, out, g0);
}

}  // namespace sz3b
