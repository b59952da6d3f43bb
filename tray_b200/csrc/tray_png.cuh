// tray_png.cuh -- SURVEY 8(f)-3: the `-save` path (SaveImage -> png.Encode, main.go:26-36, benchmark/benchmark.go:23-33)
// on the device. The frame never leaves HBM as pixels: only the finished PNG file crosses PCIe.
//
// What Go's encoder produces for an opaque *image.RGBA is an 8-bit truecolour (colour type 2) PNG; so does this one.
// Format parity is decoded-pixel equality (a PNG decoder must return exactly the rendered bytes), not byte equality
// of the file: the deflate stream here is built for a GPU -- per-row adaptive filter, then DYNAMIC HUFFMAN blocks of
// literals only (no LZ77 matches: path-traced pixels are noisy, the gain is in the entropy code of the residuals),
// bit-packed in parallel.
//
//   png_filter_kernel   one CTA per row: picks the filter (None/Sub/Up/Average/Paeth) with the smallest sum of
//                       |signed residual| (the heuristic of the PNG spec, also Go's), writes the filtered scanline,
//                       the row's Adler-32 partial sums and the byte histogram of its deflate block.
//   png_huffman_kernel  one CTA per deflate block: length-limited (15 bit) Huffman code from the histogram, canonical
//                       bit-reversed codes, the block header (code-length code, 7 bit) and the block's size in bits.
//   png_layout_kernel   exclusive scan of the block sizes, Adler-32 combine, file layout (one thread; a few hundred items).
//   png_pack_kernel     one CTA per deflate block: per-thread bit counts, CTA scan, bits OR-ed into the stream.
//   png_crc_kernel / png_finish_kernel   CRC-32 of the IDAT chunk: per-4KB raw CRCs, combined with x^(8n) mod P shifts.
#pragma once
#include <cstdint>

namespace tray {

constexpr int kPngSyms = 257;        // literals 0..255 + end-of-block; no length codes are ever used
constexpr int kPngHdrWords = 72;     // block header bit buffer (<= 17 + 19*3 + 258*7 = 1880 bits)
constexpr int kPngDataByte0 = 41;    // signature 8 + IHDR chunk 25 + IDAT length 4 + "IDAT" 4
constexpr int kCrcChunk = 4096;

struct PngPlan {
    int width, height, row_len;      // row_len = 1 + 3*width
    int rows_per_block, n_blocks;
};

struct PngBlock {
    unsigned long long data_bits;    // header + symbols + end-of-block
    unsigned long long bit_offset;   // absolute bit position in the file buffer (filled by png_layout_kernel)
    unsigned hdr_nbits;
    unsigned hdr[kPngHdrWords];
    unsigned short code[kPngSyms];   // bit-reversed canonical code
    unsigned char len[kPngSyms];
};

struct PngTotals {                   // written by png_layout_kernel / png_finish_kernel, read back by the host
    unsigned long long zlib_bytes;   // length of the IDAT payload
    unsigned long long file_bytes;
    unsigned adler;
    unsigned crc;
};

__device__ __forceinline__ int png_paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
__device__ __forceinline__ int png_abs8(int v) { v &= 255; return v < 128 ? v : 256 - v; }

// residual of one channel under filter f: x = this byte, a = left, b = up, c = up-left
__device__ __forceinline__ int png_residual(int f, int x, int a, int b, int c) {
    int pred = f == 0 ? 0 : (f == 1 ? a : (f == 2 ? b : (f == 3 ? ((a + b) >> 1) : png_paeth(a, b, c))));
    return (x - pred) & 255;
}

__global__ void __launch_bounds__(256) png_filter_kernel(const uchar4* __restrict__ img, PngPlan P, unsigned char* __restrict__ filt,
                                                         unsigned* __restrict__ hist, unsigned long long* __restrict__ row_adler) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_red[5][8];
    __shared__ int s_choice;
    const int y = blockIdx.x, tid = threadIdx.x, w = P.width;
    s_hist[tid] = 0;
    const uchar4* cur = img + (size_t)y * w;
    const uchar4* up = y > 0 ? cur - w : nullptr;
    // pass 1: cost of every filter type
    unsigned long long cost[5] = {0, 0, 0, 0, 0};
    for (int x = tid; x < w; x += 256) {
        uchar4 px = cur[x];
        uchar4 pa = x > 0 ? cur[x - 1] : make_uchar4(0, 0, 0, 0);
        uchar4 pb = up ? up[x] : make_uchar4(0, 0, 0, 0);
        uchar4 pc = (up && x > 0) ? up[x - 1] : make_uchar4(0, 0, 0, 0);
        const int X[3] = {px.x, px.y, px.z}, A[3] = {pa.x, pa.y, pa.z}, B[3] = {pb.x, pb.y, pb.z}, Cc[3] = {pc.x, pc.y, pc.z};
#pragma unroll
        for (int f = 0; f < 5; f++)
#pragma unroll
            for (int k = 0; k < 3; k++) cost[f] += png_abs8(png_residual(f, X[k], A[k], B[k], Cc[k]));
    }
#pragma unroll
    for (int f = 0; f < 5; f++) {
        unsigned long long v = cost[f];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) s_red[f][tid >> 5] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        unsigned long long bc = ~0ull;
        for (int f = 0; f < 5; f++) {
            unsigned long long v = 0;
            for (int k = 0; k < 8; k++) v += s_red[f][k];
            if (v < bc) { bc = v; best = f; }
        }
        s_choice = best;
    }
    __syncthreads();
    const int f = s_choice;
    // pass 2: write the scanline, histogram, Adler partials  (s1 = sum d_i, s2 = sum (L - i) d_i over the row's L bytes)
    unsigned char* out = filt + (size_t)y * P.row_len;
    const unsigned long long L = (unsigned long long)P.row_len;
    unsigned long long s1 = 0, s2 = 0;
    if (tid == 0) { out[0] = (unsigned char)f; atomicAdd(&s_hist[f], 1u); s1 += f; s2 += L * f; }
    for (int x = tid; x < w; x += 256) {
        uchar4 px = cur[x];
        uchar4 pa = x > 0 ? cur[x - 1] : make_uchar4(0, 0, 0, 0);
        uchar4 pb = up ? up[x] : make_uchar4(0, 0, 0, 0);
        uchar4 pc = (up && x > 0) ? up[x - 1] : make_uchar4(0, 0, 0, 0);
        const int X[3] = {px.x, px.y, px.z}, A[3] = {pa.x, pa.y, pa.z}, B[3] = {pb.x, pb.y, pb.z}, Cc[3] = {pc.x, pc.y, pc.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int r = png_residual(f, X[k], A[k], B[k], Cc[k]);
            unsigned long long i = 1ull + 3ull * x + k;
            out[i] = (unsigned char)r;
            atomicAdd(&s_hist[r], 1u);
            s1 += r; s2 += (L - i) * r;
        }
    }
    for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, off); s2 += __shfl_down_sync(0xffffffffu, s2, off); }
    __syncthreads();
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = s1; s_red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0, b = 0;
        for (int k = 0; k < 8; k++) { a += s_red[0][k]; b += s_red[1][k]; }
        row_adler[2 * y] = a % 65521ull; row_adler[2 * y + 1] = b % 65521ull;
    }
    unsigned v = s_hist[tid];
    if (v) atomicAdd(&hist[(size_t)(y / P.rows_per_block) * 256 + tid], v);
}

// ---- Huffman construction (serial parts run in thread 0; the alphabets are tiny) ------------------------------------
// Length-limited code lengths from frequencies: Huffman tree by the two-queue method over the symbols sorted by
// (freq, index), then zlib's overflow redistribution (gen_bitlen) so that the code is COMPLETE with max length
// `maxbits`. order[] = used symbols in ascending (freq, index) order, m = their number (>= 2).
__device__ void png_code_lengths(const unsigned* freq, const short* order, int m, int maxbits, unsigned char* len_out, int n_sym,
                                 unsigned long long* node_w, short* parent, unsigned char* depth) {
    for (int s = 0; s < n_sym; s++) len_out[s] = 0;
    for (int i = 0; i < m; i++) node_w[i] = freq[order[i]];
    int li = 0, ii = m, k = m;  // next leaf, next internal, next free node
    while (k < 2 * m - 1) {
        int pick[2];
        for (int t = 0; t < 2; t++) {
            bool leaf = li < m && (ii >= k || node_w[li] <= node_w[ii]);
            pick[t] = leaf ? li++ : ii++;
        }
        node_w[k] = node_w[pick[0]] + node_w[pick[1]];
        parent[pick[0]] = (short)k; parent[pick[1]] = (short)k;
        k++;
    }
    depth[2 * m - 2] = 0;
    int bl_count[32];
    for (int b = 0; b < 32; b++) bl_count[b] = 0;
    int overflow = 0;
    for (int nd = 2 * m - 3; nd >= 0; nd--) {  // root downwards: interior nodes (m..2m-2) come before the leaves (0..m-1)
        int d = depth[parent[nd]] + 1;
        if (d > maxbits) { d = maxbits; overflow++; }  // as zlib's gen_bitlen: EVERY node below the limit counts, interior ones too
        depth[nd] = (unsigned char)d;
        if (nd < m) bl_count[d]++;
    }
    while (overflow > 0) {
        int bits = maxbits - 1;
        while (bl_count[bits] == 0) bits--;
        bl_count[bits]--;
        bl_count[bits + 1] += 2;
        bl_count[maxbits]--;
        overflow -= 2;
    }
    int i = 0;  // least frequent symbols take the longest codes
    for (int bits = maxbits; bits >= 1; bits--)
        for (int c = bl_count[bits]; c > 0; c--) len_out[order[i++]] = (unsigned char)bits;
}

__device__ __forceinline__ unsigned png_bitrev(unsigned code, int len) { return __brev(code) >> (32 - len); }

struct PngBitWriter {
    unsigned* w; unsigned nbits;
    __device__ void put(unsigned v, int n) {  // n <= 16
        unsigned pos = nbits & 31, idx = nbits >> 5;
        w[idx] |= v << pos;
        if (pos + n > 32) w[idx + 1] |= v >> (32 - pos);
        nbits += n;
    }
};

__global__ void __launch_bounds__(320) png_huffman_kernel(const unsigned* __restrict__ hist, PngPlan P, PngBlock* __restrict__ blocks) {
    __shared__ unsigned s_freq[kPngSyms];
    __shared__ short s_order[kPngSyms];
    __shared__ unsigned char s_len[kPngSyms];
    __shared__ unsigned long long s_w[2 * kPngSyms];
    __shared__ short s_parent[2 * kPngSyms];
    __shared__ unsigned char s_depth[2 * kPngSyms];
    __shared__ int s_m;
    __shared__ unsigned s_next[16];
    const int b = blockIdx.x, tid = threadIdx.x;
    PngBlock& B = blocks[b];
    if (tid == 0) s_m = 0;
    if (tid < kPngSyms) s_freq[tid] = tid < 256 ? hist[(size_t)b * 256 + tid] : 1u;  // end-of-block occurs once
    __syncthreads();
    if (tid < kPngSyms && s_freq[tid] > 0) {
        int rank = 0;
        const unsigned f = s_freq[tid];
        for (int u = 0; u < kPngSyms; u++) {
            unsigned g = s_freq[u];
            if (g > 0 && (g < f || (g == f && u < tid))) rank++;
        }
        s_order[rank] = (short)tid;
        atomicAdd(&s_m, 1);
    }
    __syncthreads();
    if (tid == 0) {
        // s_m >= 2 always: end-of-block plus at least one data byte (every block holds at least one scanline)
        png_code_lengths(s_freq, s_order, s_m, 15, s_len, kPngSyms, s_w, s_parent, s_depth);
        int cnt[16];
        for (int l = 0; l < 16; l++) cnt[l] = 0;
        for (int s = 0; s < kPngSyms; s++) cnt[s_len[s]]++;
        unsigned code = 0;
        cnt[0] = 0;
        for (int l = 1; l < 16; l++) { code = (code + cnt[l - 1]) << 1; s_next[l] = code; }
    }
    __syncthreads();
    if (tid < kPngSyms) {
        const int l = s_len[tid];
        unsigned c = 0;
        if (l) {
            int before = 0;
            for (int u = 0; u < tid; u++) before += s_len[u] == l;
            c = png_bitrev(s_next[l] + before, l);
        }
        B.code[tid] = (unsigned short)c;
        B.len[tid] = (unsigned char)l;
    }
    __syncthreads();
    if (tid == 0) {
        // ---- block header: BFINAL, BTYPE=10, HLIT=257, HDIST=1, the code-length code, then the 258 lengths verbatim ----
        unsigned cl_freq[19];
        for (int s = 0; s < 19; s++) cl_freq[s] = 0;
        for (int s = 0; s < kPngSyms; s++) cl_freq[s_len[s]]++;
        cl_freq[1]++;  // the single distance code: length 1 (never used; a lone 1-bit code is a legal incomplete set)
        short cl_order[19];
        int cm = 0;
        for (int s = 0; s < 19; s++) if (cl_freq[s]) cl_order[cm++] = (short)s;
        for (int i = 1; i < cm; i++) {  // insertion sort by (freq, index)
            short v = cl_order[i];
            int j = i - 1;
            while (j >= 0 && (cl_freq[cl_order[j]] > cl_freq[v] || (cl_freq[cl_order[j]] == cl_freq[v] && cl_order[j] > v))) { cl_order[j + 1] = cl_order[j]; j--; }
            cl_order[j + 1] = v;
        }
        unsigned char cl_len[19];
        // cm >= 2 always: length 1 (distance code) plus either a 0 (unused literal) or two distinct literal lengths
        png_code_lengths(cl_freq, cl_order, cm, 7, cl_len, 19, s_w, s_parent, s_depth);
        unsigned cl_code[19];
        {
            int cnt[8];
            for (int l = 0; l < 8; l++) cnt[l] = 0;
            for (int s = 0; s < 19; s++) cnt[cl_len[s]]++;
            unsigned next[8], code = 0;
            cnt[0] = 0;
            for (int l = 1; l < 8; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
            for (int s = 0; s < 19; s++) cl_code[s] = cl_len[s] ? png_bitrev(next[cl_len[s]]++, cl_len[s]) : 0;
        }
        for (int k = 0; k < kPngHdrWords; k++) B.hdr[k] = 0;
        PngBitWriter W{B.hdr, 0};
        W.put(b == P.n_blocks - 1 ? 1u : 0u, 1);
        W.put(2u, 2);
        W.put(0u, 5);   // HLIT  = 257 - 257
        W.put(0u, 5);   // HDIST = 1 - 1
        const int perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        int hclen = 19;
        while (hclen > 4 && cl_len[perm[hclen - 1]] == 0) hclen--;
        W.put((unsigned)(hclen - 4), 4);
        for (int k = 0; k < hclen; k++) W.put(cl_len[perm[k]], 3);
        for (int s = 0; s < kPngSyms; s++) W.put(cl_code[s_len[s]], cl_len[s_len[s]]);
        W.put(cl_code[1], cl_len[1]);  // distance code 0: length 1
        B.hdr_nbits = W.nbits;
        unsigned long long bits = W.nbits;
        for (int s = 0; s < 256; s++) bits += (unsigned long long)hist[(size_t)b * 256 + s] * s_len[s];
        bits += s_len[256];
        B.data_bits = bits;
    }
}

__global__ void png_layout_kernel(PngPlan P, PngBlock* blocks, const unsigned long long* __restrict__ row_adler, PngTotals* tot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long pos = (unsigned long long)kPngDataByte0 * 8 + 16;  // after the 2-byte zlib header
    for (int b = 0; b < P.n_blocks; b++) { blocks[b].bit_offset = pos; pos += blocks[b].data_bits; }
    unsigned long long A = 1, Bv = 0;
    const unsigned long long L = (unsigned long long)P.row_len, M = 65521ull;
    for (int y = 0; y < P.height; y++) {  // Adler-32 of the concatenated scanlines from the per-row partial sums
        Bv = (Bv + (L % M) * A + row_adler[2 * y + 1]) % M;
        A = (A + row_adler[2 * y]) % M;
    }
    tot->adler = (unsigned)((Bv << 16) | A);
    unsigned long long end_byte = (pos + 7) / 8;         // deflate stream padded to a byte
    tot->zlib_bytes = end_byte - kPngDataByte0 + 4;      // + Adler-32
    tot->file_bytes = end_byte + 4 + 4 + 12;             // + Adler-32, chunk CRC, IEND chunk
}

constexpr int kPackThreads = 512, kPackBytes = 32;  // bytes per thread per tile

__global__ void __launch_bounds__(kPackThreads) png_pack_kernel(const unsigned char* __restrict__ filt, PngPlan P, const PngBlock* __restrict__ blocks,
                                                                unsigned* __restrict__ out_words) {
    __shared__ unsigned short s_code[kPngSyms];
    __shared__ unsigned char s_len[kPngSyms];
    __shared__ unsigned s_warp[kPackThreads / 32];
    __shared__ unsigned long long s_base;
    const int b = blockIdx.x, tid = threadIdx.x;
    const PngBlock& B = blocks[b];
    for (int s = tid; s < kPngSyms; s += kPackThreads) { s_code[s] = B.code[s]; s_len[s] = B.len[s]; }
    const unsigned long long bit0 = B.bit_offset;
    // header
    for (int k = tid; k * 32 < (int)B.hdr_nbits; k += kPackThreads) {
        unsigned long long p = bit0 + 32ull * k;
        unsigned v = B.hdr[k], sh = (unsigned)(p & 31);
        atomicOr(&out_words[p >> 5], v << sh);
        if (sh && (v >> (32 - sh))) atomicOr(&out_words[(p >> 5) + 1], v >> (32 - sh));
    }
    if (tid == 0) s_base = bit0 + B.hdr_nbits;
    __syncthreads();
    const size_t row0 = (size_t)b * P.rows_per_block;
    const size_t rows = min((size_t)P.rows_per_block, (size_t)P.height - row0);
    const size_t n_bytes = rows * P.row_len;
    const unsigned char* src = filt + row0 * P.row_len;
    for (size_t tile = 0; tile < n_bytes; tile += (size_t)kPackThreads * kPackBytes) {
        const size_t beg = tile + (size_t)tid * kPackBytes;
        const int n = beg < n_bytes ? (int)min((size_t)kPackBytes, n_bytes - beg) : 0;
        unsigned char v[kPackBytes];
        unsigned bits = 0;
        for (int i = 0; i < n; i++) { v[i] = src[beg + i]; bits += s_len[v[i]]; }
        // CTA exclusive scan of the bit counts
        unsigned incl = bits;
        for (int off = 1; off < 32; off <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, off); if ((tid & 31) >= off) incl += t; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        unsigned warp_off = 0, total = 0;
        for (int k = 0; k < kPackThreads / 32; k++) { unsigned t = s_warp[k]; if (k < (tid >> 5)) warp_off += t; total += t; }
        unsigned long long p = s_base + warp_off + (incl - bits);
        // emit: 64-bit accumulator, 32-bit words OR-ed into the (zeroed) stream
        unsigned long long acc = 0;
        unsigned nb = (unsigned)(p & 31);
        size_t wi = (size_t)(p >> 5);
        for (int i = 0; i < n; i++) {
            acc |= (unsigned long long)s_code[v[i]] << nb;
            nb += s_len[v[i]];
            if (nb >= 32) { atomicOr(&out_words[wi++], (unsigned)acc); acc >>= 32; nb -= 32; }
        }
        if (nb > 0 && (unsigned)acc) atomicOr(&out_words[wi], (unsigned)acc);
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) {  // end-of-block
        unsigned long long p = s_base;
        unsigned v = s_code[256], sh = (unsigned)(p & 31);
        atomicOr(&out_words[p >> 5], v << sh);
        if (sh + s_len[256] > 32) atomicOr(&out_words[(p >> 5) + 1], v >> (32 - sh));
    }
}

// ---- CRC-32 (reflected, polynomial 0xEDB88320) ----------------------------------------------------------------------
__device__ __forceinline__ unsigned crc_table_entry(unsigned i) {
    unsigned c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    return c;
}
// product of two polynomials mod P in the reflected representation (bit 31 = x^0)
__device__ __forceinline__ unsigned crc_mulmod(unsigned a, unsigned b) {
    unsigned p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// x^(8n) mod P
__device__ unsigned crc_xpow8n(unsigned long long n) {
    unsigned r = 0x80000000u;       // 1
    unsigned sq = 0x00800000u;      // x^8
    while (n) {
        if (n & 1) r = crc_mulmod(r, sq);
        sq = crc_mulmod(sq, sq);
        n >>= 1;
    }
    return r;
}

// raw CRC (init 0, no final xor) of every 4 KB piece of bytes [first, last)
__global__ void __launch_bounds__(256) png_crc_kernel(const unsigned char* __restrict__ buf, const PngTotals* __restrict__ tot,
                                                      unsigned* __restrict__ piece_crc, unsigned long long cap_pieces) {
    __shared__ unsigned s_tab[256];
    s_tab[threadIdx.x] = crc_table_entry(threadIdx.x);
    __syncthreads();
    const unsigned long long first = kPngDataByte0 - 4, last = kPngDataByte0 + tot->zlib_bytes;  // "IDAT" + payload
    const unsigned long long piece = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long beg = first + piece * kCrcChunk;
    if (beg >= last || piece >= cap_pieces) return;
    const unsigned long long end = beg + kCrcChunk < last ? beg + kCrcChunk : last;
    unsigned c = 0;
    for (unsigned long long i = beg; i < end; i++) c = s_tab[(c ^ buf[i]) & 255] ^ (c >> 8);
    piece_crc[piece] = c;
}

// writes the Adler-32 (before png_crc_kernel runs, mode 0) or combines the piece CRCs and writes CRC + IEND (mode 1)
__global__ void __launch_bounds__(256) png_finish_kernel(unsigned char* __restrict__ buf, PngTotals* tot, const unsigned* __restrict__ piece_crc, int mode) {
    __shared__ unsigned s_crc[256], s_mul[256];
    const int tid = threadIdx.x;
    const unsigned long long zend = kPngDataByte0 + tot->zlib_bytes;
    if (mode == 0) {
        if (tid == 0) {
            unsigned a = tot->adler;
            buf[zend - 4] = (unsigned char)(a >> 24); buf[zend - 3] = (unsigned char)(a >> 16); buf[zend - 2] = (unsigned char)(a >> 8); buf[zend - 1] = (unsigned char)a;
            unsigned long long n = tot->zlib_bytes;
            buf[kPngDataByte0 - 8] = (unsigned char)(n >> 24); buf[kPngDataByte0 - 7] = (unsigned char)(n >> 16);
            buf[kPngDataByte0 - 6] = (unsigned char)(n >> 8); buf[kPngDataByte0 - 5] = (unsigned char)n;
        }
        return;
    }
    const unsigned long long total = tot->zlib_bytes + 4;  // bytes covered by the chunk CRC
    const unsigned long long n_pieces = (total + kCrcChunk - 1) / kCrcChunk;
    const unsigned long long per = (n_pieces + 255) / 256;
    // thread t folds pieces [t*per, (t+1)*per): acc = acc * x^(8*len_piece) + crc_piece
    unsigned acc = 0;
    unsigned long long len = 0;
    const unsigned xfull = crc_xpow8n(kCrcChunk);
    for (unsigned long long k = (unsigned long long)tid * per; k < (unsigned long long)(tid + 1) * per && k < n_pieces; k++) {
        unsigned long long plen = (k + 1) * kCrcChunk <= total ? kCrcChunk : total - k * kCrcChunk;
        acc = crc_mulmod(acc, plen == kCrcChunk ? xfull : crc_xpow8n(plen)) ^ piece_crc[k];
        len += plen;
    }
    s_crc[tid] = acc;
    s_mul[tid] = crc_xpow8n(len);
    __syncthreads();
    if (tid == 0) {
        unsigned c = 0;
        for (int t = 0; t < 256; t++) c = crc_mulmod(c, s_mul[t]) ^ s_crc[t];
        // standard CRC = raw CRC with the 0xFFFFFFFF preset pushed through the whole message, final complement
        c ^= crc_mulmod(0xFFFFFFFFu, crc_xpow8n(total)) ^ 0xFFFFFFFFu;
        tot->crc = c;
        unsigned char* q = buf + zend;
        q[0] = (unsigned char)(c >> 24); q[1] = (unsigned char)(c >> 16); q[2] = (unsigned char)(c >> 8); q[3] = (unsigned char)c;
        const unsigned char iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
        for (int k = 0; k < 12; k++) q[4 + k] = iend[k];
    }
}

}  // namespace tray
