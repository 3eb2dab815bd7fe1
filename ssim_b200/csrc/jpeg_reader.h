// jpeg_reader.h -- JPEG reader of the `rmgr-ssim` front end (baseline and progressive Huffman, 8 bits per sample).
//
// Why it exists: the reference decodes its inputs with stb_image (src/ssim-cli.cpp:33-40, tests/rmgr-ssim-tests.cpp:237-243),
// which it downloads at configure time (CMakeLists.txt:242-261) and which is not available offline.  22 of the reference's
// 30 test images are JPEGs, and the 132 known answers of its bbb suites (tests/rmgr-ssim-tests.cpp:388-465) are SSIMs of
// the pixels THAT decoder produces.  A JPEG's coefficients are defined exactly by the bit stream; what differs between
// decoders is the arithmetic after them.  This reader therefore follows the published arithmetic of stb_image's scalar path
// (which its SIMD paths reproduce bit for bit):
//   * inverse DCT: the 12-bit fixed-point Loeffler form of the IJG "islow" transform with 2 extra bits kept between the
//     column and the row pass (rounding constants 512 / 65536, level shift folded into the row pass),
//   * YCbCr -> RGB in 20-bit fixed point with the chroma-to-green term truncated to its upper 16 bits,
//   * chroma upsampling by the 3:1 triangle filters (h2, v2, h2v2), nearest neighbour otherwise.
// tests/test_jpeg.py pins it: the reference's own expected SSIMs are reproduced to 1e-13 from the reference's JPEG files,
// which only a decoder with identical pixels can do.  Written from the JPEG standard (ITU T.81) and the description above.
#ifndef SSIM_B200_JPEG_READER_H
#define SSIM_B200_JPEG_READER_H

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace jpegr
{

struct Huffman
{
    // canonical code tables (T.81 annex C / F.2.2.3): codes of length L are the values [first[L], first[L] + count[L])
    int      count[17];
    int      firstCode[17];
    int      firstIndex[17];
    uint8_t  values[256];
    uint8_t  lookLen[512];      // 9-bit prefix -> code length (0 = longer than 9 bits)
    uint8_t  lookVal[512];
    bool     present;
    Huffman() : present(false) {}

    bool build(const uint8_t counts[16], const uint8_t* vals, int n)
    {
        int code = 0, index = 0;
        for (int len = 1; len <= 16; ++len) {
            count[len] = counts[len - 1];
            firstCode[len] = code;
            firstIndex[len] = index;
            code += count[len];
            index += count[len];
            if (code > (1 << len)) return false;
            code <<= 1;
        }
        if (index != n || n > 256) return false;
        memcpy(values, vals, (size_t)n);
        memset(lookLen, 0, sizeof(lookLen));
        for (int len = 1; len <= 9; ++len)
            for (int i = 0; i < count[len]; ++i) {
                const int c = (firstCode[len] + i) << (9 - len);
                for (int f = 0; f < (1 << (9 - len)); ++f) { lookLen[c + f] = (uint8_t)len; lookVal[c + f] = values[firstIndex[len] + i]; }
            }
        present = true;
        return true;
    }
};

struct Component
{
    int id, h, v, tq;
    int td, ta;                 // tables of the current scan
    int x, y;                   // size in samples
    int w2, h2;                 // size padded to whole MCUs
    int dcPred;
    std::vector<uint8_t> plane; // w2 x h2 samples
    std::vector<short>   coeff; // progressive: (w2/8) x (h2/8) blocks of 64
    int coeffW;
};

static const uint8_t kZigzag[64 + 15] = {
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };   // + 15 so that a corrupt run cannot index out of the block

class Decoder
{
public:
    std::string error;
    int width, height, channels;
    std::vector<uint8_t> pixels;    // interleaved, top-down

    Decoder() : width(0), height(0), channels(0), d_(0), n_(0), pos_(0) {}

    bool decode(const uint8_t* data, size_t size)
    {
        d_ = data; n_ = size; pos_ = 0;
        progressive_ = false; restartInterval_ = 0; haveFrame_ = false; jfif_ = false; adobeTransform_ = -1;
        pendingMarker_ = -1; memset(quant_, 0, sizeof(quant_)); memset(haveQuant_, 0, sizeof(haveQuant_));
        if (n_ < 4 || d_[0] != 0xFF || d_[1] != 0xD8) return fail("not a JPEG");
        pos_ = 2;
        for (;;) {
            int m = next_marker();
            if (m < 0) return fail("truncated JPEG (no EOI)");
            if (m == 0xD9) break;                                   // EOI
            if (m == 0xDA) {                                        // SOS + entropy-coded data
                if (!read_scan_header() || !decode_scan()) return false;
                continue;
            }
            if (!read_segment(m)) return false;
        }
        if (!haveFrame_) return fail("no frame header");
        for (size_t k = 0; k < comps_.size(); ++k) if (!haveQuant_[comps_[k].tq]) return fail("missing quantisation table");
        if (progressive_) finish_progressive();
        return convert();
    }

private:
    const uint8_t* d_;
    size_t n_, pos_;
    bool progressive_, haveFrame_, jfif_;
    int adobeTransform_;
    int restartInterval_;
    uint16_t quant_[4][64];      // natural order
    bool haveQuant_[4];
    Huffman dc_[4], ac_[4];
    std::vector<Component> comps_;
    int hMax_, vMax_, mcuX_, mcuY_;
    // scan state
    int scanN_, scanOrder_[4], ss_, se_, ah_, al_, eobRun_;
    uint32_t bitBuf_; int bitCnt_; int pendingMarker_; bool noMore_;

    bool fail(const char* msg) { error = msg; return false; }
    int  u8()  { return pos_ < n_ ? d_[pos_++] : 0; }
    int  u16() { const int a = u8(); return (a << 8) | u8(); }

    // next marker at or after pos_ (skips fill bytes and stray data, as decoders conventionally do)
    int next_marker()
    {
        while (pos_ < n_) {
            if (d_[pos_++] != 0xFF) continue;
            while (pos_ < n_ && d_[pos_] == 0xFF) ++pos_;
            if (pos_ >= n_) return -1;
            const int m = d_[pos_++];
            if (m != 0x00) return m;
        }
        return -1;
    }

    bool read_segment(int m)
    {
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) return true;     // TEM, stray RSTn: no payload
        if (pos_ + 2 > n_) return fail("truncated JPEG segment");
        int len = u16();
        if (len < 2 || pos_ + (size_t)(len - 2) > n_) return fail("bad JPEG segment length");
        const size_t end = pos_ + (size_t)(len - 2);
        switch (m) {
            case 0xDB:                                              // DQT
                while (pos_ < end) {
                    const int pq = u8(), prec = pq >> 4, t = pq & 15;
                    if (prec > 1 || t > 3) return fail("bad DQT");
                    for (int i = 0; i < 64; ++i) quant_[t][kZigzag[i]] = (uint16_t)(prec ? u16() : u8());
                    haveQuant_[t] = true;
                }
                break;
            case 0xC4:                                              // DHT
                while (pos_ < end) {
                    const int tc = u8(), cls = tc >> 4, t = tc & 15;
                    if (cls > 1 || t > 3 || pos_ + 16 > end) return fail("bad DHT");
                    uint8_t counts[16]; int total = 0;
                    for (int i = 0; i < 16; ++i) { counts[i] = (uint8_t)u8(); total += counts[i]; }
                    if (total > 256 || pos_ + (size_t)total > end) return fail("bad DHT");
                    if (!(cls ? ac_[t] : dc_[t]).build(counts, d_ + pos_, total)) return fail("bad Huffman code lengths");
                    pos_ += (size_t)total;
                }
                break;
            case 0xDD: restartInterval_ = u16(); break;             // DRI
            case 0xC0: case 0xC1: case 0xC2:                        // SOF0/1/2
                if (haveFrame_) return fail("multiple frame headers");
                progressive_ = (m == 0xC2);
                if (!read_frame(end)) return false;
                break;
            case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                return fail("unsupported JPEG coding process (lossless / hierarchical / arithmetic)");
            case 0xE0:                                              // APP0: JFIF
                if (len >= 7 && !memcmp(d_ + pos_, "JFIF\0", 5)) jfif_ = true;
                break;
            case 0xEE:                                              // APP14: Adobe colour transform flag
                if (len >= 14 && !memcmp(d_ + pos_, "Adobe\0", 6)) adobeTransform_ = d_[pos_ + 11];
                break;
            default: break;                                         // other APPn, COM, ...: skipped
        }
        pos_ = end;
        return true;
    }

    bool read_frame(size_t end)
    {
        if (end - pos_ < 6) return fail("bad SOF");
        if (u8() != 8) return fail("only 8-bit JPEG is supported");
        height = u16(); width = u16();
        const int nc = u8();
        if (width <= 0 || height <= 0) return fail("bad JPEG dimensions");
        if ((uint64_t)width * (uint64_t)height > (1ull << 28)) return fail("JPEG larger than 2^28 pixels");
        if (nc != 1 && nc != 3) return fail("unsupported JPEG component count (1 or 3 expected)");
        if (end - pos_ < (size_t)(3 * nc)) return fail("bad SOF");
        comps_.assign((size_t)nc, Component());
        hMax_ = vMax_ = 1;
        for (int i = 0; i < nc; ++i) {
            Component& c = comps_[(size_t)i];
            c.id = u8();
            const int hv = u8();
            c.h = hv >> 4; c.v = hv & 15; c.tq = u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad SOF component");
            if (c.h > hMax_) hMax_ = c.h;
            if (c.v > vMax_) vMax_ = c.v;
        }
        for (int i = 0; i < nc; ++i) if (hMax_ % comps_[(size_t)i].h || vMax_ % comps_[(size_t)i].v) return fail("unsupported sampling factors");
        mcuX_ = (width  + 8 * hMax_ - 1) / (8 * hMax_);
        mcuY_ = (height + 8 * vMax_ - 1) / (8 * vMax_);
        for (int i = 0; i < nc; ++i) {
            Component& c = comps_[(size_t)i];
            c.x  = (width  * c.h + hMax_ - 1) / hMax_;
            c.y  = (height * c.v + vMax_ - 1) / vMax_;
            c.w2 = mcuX_ * c.h * 8;
            c.h2 = mcuY_ * c.v * 8;
            c.plane.assign((size_t)c.w2 * c.h2, 0);
            c.coeffW = c.w2 / 8;
            if (progressive_) c.coeff.assign((size_t)c.w2 * c.h2, 0);
            c.dcPred = 0; c.td = c.ta = 0;
        }
        haveFrame_ = true;
        return true;
    }

    bool read_scan_header()
    {
        if (!haveFrame_) return fail("scan before frame header");
        const int len = u16();
        scanN_ = u8();
        if (scanN_ < 1 || scanN_ > (int)comps_.size() || len != 6 + 2 * scanN_) return fail("bad SOS");
        for (int i = 0; i < scanN_; ++i) {
            const int id = u8(), t = u8();
            int which = -1;
            for (size_t k = 0; k < comps_.size(); ++k) if (comps_[k].id == id) which = (int)k;
            if (which < 0) return fail("bad SOS component");
            comps_[(size_t)which].td = t >> 4; comps_[(size_t)which].ta = t & 15;
            if (comps_[(size_t)which].td > 3 || comps_[(size_t)which].ta > 3) return fail("bad SOS table index");
            scanOrder_[i] = which;
        }
        ss_ = u8(); se_ = u8();
        const int a = u8();
        ah_ = a >> 4; al_ = a & 15;
        if (progressive_) {
            if (ss_ > 63 || se_ > 63 || ss_ > se_ || ah_ > 13 || al_ > 13) return fail("bad progressive SOS");
            if (ss_ == 0 && se_ != 0) return fail("progressive scan mixes DC and AC");
            if (ss_ > 0 && scanN_ != 1) return fail("interleaved progressive AC scan");
        } else {
            if (ss_ != 0 || ah_ != 0 || al_ != 0) return fail("bad baseline SOS");
            se_ = 63;
        }
        return true;
    }

    // ---- entropy-coded segment: MSB-first bits, FF00 unstuffing, a marker ends the feed (zeros afterwards)
    void reset_bits() { bitBuf_ = 0; bitCnt_ = 0; pendingMarker_ = -1; noMore_ = false; eobRun_ = 0; for (size_t k = 0; k < comps_.size(); ++k) comps_[k].dcPred = 0; }
    void fill()
    {
        while (bitCnt_ <= 24) {
            int b = 0;
            if (!noMore_) {
                if (pos_ >= n_) noMore_ = true;
                else {
                    b = d_[pos_++];
                    if (b == 0xFF) {
                        int c = pos_ < n_ ? d_[pos_] : 0xD9;
                        while (c == 0xFF && pos_ + 1 < n_) { ++pos_; c = d_[pos_]; }   // fill bytes
                        if (c == 0x00) ++pos_;
                        else { pendingMarker_ = c; --pos_; noMore_ = true; b = 0; }   // leave pos_ on the FF of the marker
                    }
                }
            }
            bitBuf_ |= (uint32_t)b << (24 - bitCnt_);
            bitCnt_ += 8;
        }
    }
    int get_bits(int n)
    {
        if (n == 0) return 0;
        if (bitCnt_ < n) fill();
        const int v = (int)(bitBuf_ >> (32 - n));
        bitBuf_ <<= n; bitCnt_ -= n;
        return v;
    }
    int get_bit() { return get_bits(1); }
    int extend_receive(int n)                                       // T.81 F.2.2.1 RECEIVE + EXTEND
    {
        if (n == 0) return 0;
        const int v = get_bits(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    int decode_symbol(const Huffman& h)
    {
        if (bitCnt_ < 16) fill();
        const int look = (int)(bitBuf_ >> 23);
        if (h.lookLen[look]) { const int len = h.lookLen[look]; bitBuf_ <<= len; bitCnt_ -= len; return h.lookVal[look]; }
        int code = (int)(bitBuf_ >> 22);                            // 10 bits
        for (int len = 10; len <= 16; ++len) {
            if (code - h.firstCode[len] < h.count[len] && code >= h.firstCode[len]) {
                bitBuf_ <<= len; bitCnt_ -= len;
                return h.values[h.firstIndex[len] + code - h.firstCode[len]];
            }
            code = (int)(bitBuf_ >> (31 - len));                    // one more bit
        }
        return -1;
    }

    bool block_baseline(Component& c, short* blk)
    {
        memset(blk, 0, 64 * sizeof(short));
        const Huffman &hd = dc_[c.td], &ha = ac_[c.ta];
        const int t = decode_symbol(hd);
        if (t < 0 || t > 15) return fail("bad Huffman code");
        c.dcPred = (int)((uint32_t)c.dcPred + (uint32_t)extend_receive(t));
        blk[0] = (short)((uint32_t)c.dcPred * quant_[c.tq][0]);
        for (int k = 1; k < 64;) {
            const int rs = decode_symbol(ha);
            if (rs < 0) return fail("bad Huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) { if (rs != 0xF0) break; k += 16; }
            else { k += r; const int z = kZigzag[k++]; blk[z] = (short)((uint32_t)extend_receive(s) * quant_[c.tq][z]); }
        }
        return true;
    }
    bool block_prog_dc(Component& c, short* blk)
    {
        if (ah_ == 0) {
            memset(blk, 0, 64 * sizeof(short));
            const int t = decode_symbol(dc_[c.td]);
            if (t < 0 || t > 15) return fail("bad Huffman code");
            c.dcPred = (int)((uint32_t)c.dcPred + (uint32_t)extend_receive(t));
            blk[0] = (short)((uint32_t)c.dcPred << al_);
        } else if (get_bit()) blk[0] = (short)(blk[0] + (1 << al_));
        return true;
    }
    bool block_prog_ac(Component& c, short* blk)
    {
        const Huffman& ha = ac_[c.ta];
        if (ah_ == 0) {                                             // first pass of this band (T.81 G.1.2.2)
            if (eobRun_) { --eobRun_; return true; }
            int k = ss_;
            do {
                const int rs = decode_symbol(ha);
                if (rs < 0) return fail("bad Huffman code");
                const int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) { eobRun_ = (1 << r); if (r) eobRun_ += get_bits(r); --eobRun_; break; }
                    k += 16;
                } else { k += r; blk[kZigzag[k++]] = (short)(extend_receive(s) * (1 << al_)); }
            } while (k <= se_);
        } else {                                                    // refinement (T.81 G.1.2.3)
            const short bit = (short)(1 << al_);
            if (eobRun_) {
                --eobRun_;
                for (int k = ss_; k <= se_; ++k) {
                    short* p = &blk[kZigzag[k]];
                    if (*p != 0 && get_bit() && (*p & bit) == 0) *p = (short)(*p + (*p > 0 ? bit : -bit));
                }
            } else {
                int k = ss_;
                do {
                    const int rs = decode_symbol(ha);
                    if (rs < 0) return fail("bad Huffman code");
                    int s = rs & 15, r = rs >> 4;
                    if (s == 0) {
                        if (r < 15) { eobRun_ = (1 << r) - 1; if (r) eobRun_ += get_bits(r); r = 64; }   // rest of the block: corrections only
                    } else {
                        if (s != 1) return fail("bad Huffman code");
                        s = get_bit() ? bit : -bit;
                    }
                    while (k <= se_) {
                        short* p = &blk[kZigzag[k++]];
                        if (*p != 0) { if (get_bit() && (*p & bit) == 0) *p = (short)(*p + (*p > 0 ? bit : -bit)); }
                        else { if (r == 0) { *p = (short)s; break; } --r; }
                    }
                } while (k <= se_);
            }
        }
        return true;
    }

    bool restart_if_due(int& todo)
    {
        if (--todo > 0) return true;
        // a restart marker is expected here; without one the scan simply ends
        if (bitCnt_ < 24) fill();
        if (!(pendingMarker_ >= 0xD0 && pendingMarker_ <= 0xD7)) return false;
        pos_ += 2;                                                  // consume FF Dn
        reset_bits();
        todo = restartInterval_ ? restartInterval_ : 0x7fffffff;
        return true;
    }

    bool decode_scan()
    {
        reset_bits();
        for (int i = 0; i < scanN_; ++i) {
            const Component& c = comps_[(size_t)scanOrder_[i]];
            if ((!progressive_ || (ss_ == 0 && ah_ == 0)) && !dc_[c.td].present) return fail("missing DC Huffman table");
            if ((!progressive_ || ss_ > 0) && !ac_[c.ta].present) return fail("missing AC Huffman table");
        }
        int todo = restartInterval_ ? restartInterval_ : 0x7fffffff;
        short blk[64];
        bool more = true;
        if (scanN_ == 1) {
            // non-interleaved: the component's own blocks in raster order, only those that hold image samples (T.81 A.2.3)
            Component& c = comps_[(size_t)scanOrder_[0]];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int by = 0; by < bh && more; ++by)
                for (int bx = 0; bx < bw && more; ++bx) {
                    if (progressive_) {
                        short* b = &c.coeff[64 * ((size_t)by * c.coeffW + bx)];
                        if (!(ss_ == 0 ? block_prog_dc(c, b) : block_prog_ac(c, b))) return false;
                    } else {
                        if (!block_baseline(c, blk)) return false;
                        idct(&c.plane[(size_t)by * 8 * c.w2 + (size_t)bx * 8], c.w2, blk);
                    }
                    more = restart_if_due(todo);
                }
        } else {
            for (int my = 0; my < mcuY_ && more; ++my)
                for (int mx = 0; mx < mcuX_ && more; ++mx) {
                    for (int i = 0; i < scanN_; ++i) {
                        Component& c = comps_[(size_t)scanOrder_[i]];
                        for (int v = 0; v < c.v; ++v)
                            for (int h = 0; h < c.h; ++h) {
                                const int bx = mx * c.h + h, by = my * c.v + v;
                                if (progressive_) {
                                    if (!block_prog_dc(c, &c.coeff[64 * ((size_t)by * c.coeffW + bx)])) return false;
                                } else {
                                    if (!block_baseline(c, blk)) return false;
                                    idct(&c.plane[(size_t)by * 8 * c.w2 + (size_t)bx * 8], c.w2, blk);
                                }
                            }
                    }
                    more = restart_if_due(todo);
                }
        }
        // hand the position back to the marker parser: pos_ is behind the last byte fed into the bit buffer, or on the FF
        // of the marker that ended the feed
        return true;
    }

    void finish_progressive()
    {
        for (size_t k = 0; k < comps_.size(); ++k) {
            Component& c = comps_[k];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int by = 0; by < bh; ++by)
                for (int bx = 0; bx < bw; ++bx) {
                    short* b = &c.coeff[64 * ((size_t)by * c.coeffW + bx)];
                    for (int i = 0; i < 64; ++i) b[i] = (short)((uint32_t)(int)b[i] * quant_[c.tq][i]);
                    idct(&c.plane[(size_t)by * 8 * c.w2 + (size_t)bx * 8], c.w2, b);
                }
        }
    }

    // ---- inverse DCT: 12-bit fixed-point Loeffler/IJG "islow" butterflies; the column pass keeps 2 extra bits
    static inline int fx(float x) { return (int)(x * 4096 + 0.5); }     // float constant, sum in double, truncation (also for negatives)
    static inline uint8_t clamp8(int x) { return (unsigned)x > 255 ? (x < 0 ? 0 : 255) : (uint8_t)x; }
    // 32-bit two's-complement arithmetic that wraps instead of overflowing: valid streams never get near the limit, corrupt ones
    // (coefficients of +-32767 under 16-bit quantisers) do, and a file must not be able to trigger undefined behaviour
    struct W {
        uint32_t v;
        W() : v(0) {}
        W(int x) : v((uint32_t)x) {}
        static W raw(uint32_t u) { W w; w.v = u; return w; }
        friend W operator+(W a, W b) { return raw(a.v + b.v); }
        friend W operator-(W a, W b) { return raw(a.v - b.v); }
        friend W operator*(W a, W b) { return raw(a.v * b.v); }
        int shr(int n) const { return (int)(int32_t)v >> n; }           // arithmetic shift of the signed value
    };
    struct Idct1D { W x0, x1, x2, x3, t0, t1, t2, t3; };
    static inline void idct_1d(W s0, W s1, W s2, W s3, W s4, W s5, W s6, W s7, Idct1D& o)
    {
        static const int c0541 = fx(0.5411961f), c1847 = fx(-1.847759065f), c0765 = fx(0.765366865f), c1175 = fx(1.175875602f),
                         c0298 = fx(0.298631336f), c2053 = fx(2.053119869f), c3072 = fx(3.072711026f), c1501 = fx(1.501321110f),
                         c0899 = fx(-0.899976223f), c2562 = fx(-2.562915447f), c1961 = fx(-1.961570560f), c0390 = fx(-0.390180644f);
        W p1 = (s2 + s6) * W(c0541);
        W t2 = p1 + s6 * W(c1847);
        W t3 = p1 + s2 * W(c0765);
        W t0 = (s0 + s4) * W(4096);
        W t1 = (s0 - s4) * W(4096);
        o.x0 = t0 + t3; o.x3 = t0 - t3; o.x1 = t1 + t2; o.x2 = t1 - t2;
        t0 = s7; t1 = s5; t2 = s3; t3 = s1;
        W p3 = t0 + t2, p4 = t1 + t3;
        p1 = t0 + t3;
        W p2 = t1 + t2;
        const W p5 = (p3 + p4) * W(c1175);
        t0 = t0 * W(c0298); t1 = t1 * W(c2053); t2 = t2 * W(c3072); t3 = t3 * W(c1501);
        p1 = p5 + p1 * W(c0899);
        p2 = p5 + p2 * W(c2562);
        p3 = p3 * W(c1961);
        p4 = p4 * W(c0390);
        o.t3 = t3 + p1 + p4; o.t2 = t2 + p2 + p3; o.t1 = t1 + p2 + p4; o.t0 = t0 + p1 + p3;
    }
    static void idct(uint8_t* out, int stride, const short* d)
    {
        int val[64];
        Idct1D r;
        for (int i = 0; i < 8; ++i) {                               // columns
            idct_1d(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i], r);
            const W h(512);
            r.x0 = r.x0 + h; r.x1 = r.x1 + h; r.x2 = r.x2 + h; r.x3 = r.x3 + h;
            val[i]      = (r.x0 + r.t3).shr(10); val[56 + i] = (r.x0 - r.t3).shr(10);
            val[8 + i]  = (r.x1 + r.t2).shr(10); val[48 + i] = (r.x1 - r.t2).shr(10);
            val[16 + i] = (r.x2 + r.t1).shr(10); val[40 + i] = (r.x2 - r.t1).shr(10);
            val[24 + i] = (r.x3 + r.t0).shr(10); val[32 + i] = (r.x3 - r.t0).shr(10);
        }
        for (int i = 0; i < 8; ++i, out += stride) {                // rows; 1 << 17 to remove, + 128 level shift
            const int* v = val + 8 * i;
            idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], r);
            const W bias(65536 + (128 << 17));
            r.x0 = r.x0 + bias; r.x1 = r.x1 + bias; r.x2 = r.x2 + bias; r.x3 = r.x3 + bias;
            out[0] = clamp8((r.x0 + r.t3).shr(17)); out[7] = clamp8((r.x0 - r.t3).shr(17));
            out[1] = clamp8((r.x1 + r.t2).shr(17)); out[6] = clamp8((r.x1 - r.t2).shr(17));
            out[2] = clamp8((r.x2 + r.t1).shr(17)); out[5] = clamp8((r.x2 - r.t1).shr(17));
            out[3] = clamp8((r.x3 + r.t0).shr(17)); out[4] = clamp8((r.x3 - r.t0).shr(17));
        }
    }

    // ---- chroma upsampling (one output row from the two nearest input rows)
    static void up_row(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs)
    {
        if (hs == 1 && vs == 1) { memcpy(out, nearRow, (size_t)w); return; }
        if (hs == 1 && vs == 2) { for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * nearRow[i] + farRow[i] + 2) >> 2); return; }
        if (hs == 2 && vs == 1) {
            const uint8_t* in = nearRow;
            if (w == 1) { out[0] = out[1] = in[0]; return; }
            out[0] = in[0];
            out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
            int i;
            for (i = 1; i < w - 1; ++i) {
                const int n = 3 * in[i] + 2;
                out[2 * i]     = (uint8_t)((n + in[i - 1]) >> 2);
                out[2 * i + 1] = (uint8_t)((n + in[i + 1]) >> 2);
            }
            out[2 * i]     = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
            out[2 * i + 1] = in[w - 1];
            return;
        }
        if (hs == 2 && vs == 2) {
            int t1 = 3 * nearRow[0] + farRow[0];
            if (w == 1) { out[0] = out[1] = (uint8_t)((t1 + 2) >> 2); return; }
            out[0] = (uint8_t)((t1 + 2) >> 2);
            for (int i = 1; i < w; ++i) {
                const int t0 = t1;
                t1 = 3 * nearRow[i] + farRow[i];
                out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
                out[2 * i]     = (uint8_t)((3 * t1 + t0 + 8) >> 4);
            }
            out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
            return;
        }
        for (int i = 0; i < w; ++i) for (int j = 0; j < hs; ++j) out[i * hs + j] = nearRow[i];   // other ratios: nearest
    }

    bool convert()
    {
        const int nc = (int)comps_.size();
        channels = nc;
        pixels.assign((size_t)width * height * nc, 0);
        // is the 3-component data RGB already?  (component ids 'R','G','B', or an Adobe marker with transform 0)
        const bool rgb = nc == 3 && ((comps_[0].id == 'R' && comps_[1].id == 'G' && comps_[2].id == 'B') || (adobeTransform_ == 0 && !jfif_));
        struct Up { int hs, vs, ystep, ypos, wLores; const uint8_t *line0, *line1; std::vector<uint8_t> buf; };
        std::vector<Up> up((size_t)nc);
        for (int k = 0; k < nc; ++k) {
            Up& u = up[(size_t)k];
            u.hs = hMax_ / comps_[(size_t)k].h; u.vs = vMax_ / comps_[(size_t)k].v;
            u.ystep = u.vs >> 1; u.ypos = 0;
            u.wLores = (width + u.hs - 1) / u.hs;
            u.line0 = u.line1 = comps_[(size_t)k].plane.data();
            u.buf.assign((size_t)width + 2 * (size_t)u.hs + 8, 0);
        }
        const uint8_t* rows[3] = {0, 0, 0};
        for (int j = 0; j < height; ++j) {
            for (int k = 0; k < nc; ++k) {
                Up& u = up[(size_t)k];
                const Component& c = comps_[(size_t)k];
                const bool bottom = u.ystep >= (u.vs >> 1);
                if (u.hs == 1 && u.vs == 1) rows[k] = bottom ? u.line1 : u.line0;
                else { up_row(u.buf.data(), bottom ? u.line1 : u.line0, bottom ? u.line0 : u.line1, u.wLores, u.hs, u.vs); rows[k] = u.buf.data(); }
                if (++u.ystep >= u.vs) {
                    u.ystep = 0;
                    u.line0 = u.line1;
                    if (++u.ypos < c.y) u.line1 += c.w2;
                }
            }
            uint8_t* o = &pixels[(size_t)j * width * nc];
            if (nc == 1) memcpy(o, rows[0], (size_t)width);
            else if (rgb) for (int i = 0; i < width; ++i) { o[3 * i] = rows[0][i]; o[3 * i + 1] = rows[1][i]; o[3 * i + 2] = rows[2][i]; }
            else {
                // 20-bit fixed point; the Cb -> G term keeps only its upper 16 bits (what a 16-bit SIMD multiply-high yields,
                // so that scalar and SIMD decoders agree bit for bit)
                const int kCrR = fx20(1.40200f), kCrG = fx20(0.71414f), kCbG = fx20(0.34414f), kCbB = fx20(1.77200f);
                for (int i = 0; i < width; ++i) {
                    const int yf = (rows[0][i] << 20) + (1 << 19);
                    const int cb = rows[1][i] - 128, cr = rows[2][i] - 128;
                    const int r = (yf + cr * kCrR) >> 20;
                    const int g = (int)(yf + cr * -kCrG + (int)((unsigned)(cb * -kCbG) & 0xffff0000u)) >> 20;
                    const int b = (yf + cb * kCbB) >> 20;
                    o[3 * i] = clamp8(r); o[3 * i + 1] = clamp8(g); o[3 * i + 2] = clamp8(b);
                }
            }
        }
        return true;
    }
    static inline int fx20(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
};

} // namespace jpegr

#endif // SSIM_B200_JPEG_READER_H
