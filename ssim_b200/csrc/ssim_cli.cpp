// ssim_cli.cpp -- `rmgr-ssim`: command-line front end of the GPU SSIM engine.
//
// Same usage, options and output as the reference's CLI (reference src/ssim-cli.cpp:73-84,130-213,216-389):
//     rmgr-ssim [options] img1 img2 [map]
//       -#  compute SSIM only for channel #           (# = 0..3)
//       -y  compute SSIM on luminance (BT.601; images with <= 2 channels: channel 0)
// prints "% 7.4f" for a single value, or "Channel n: % 7.4f" per channel plus "Average  : % 7.4f".
// The optional map is written as .pfm (raw floats, bottom-up) or as 8 bits per sample (max(0,s)*255) in
// .png / .bmp / .tga / .pgm / .ppm.
//
// The reference decodes images with stb_image, which it downloads at configure time and which is not available
// offline; this front end carries its own small readers instead: binary PNM (P5/P6, maxval 255), PNG (8-bit
// gray, gray+alpha, RGB, RGBA, palette; non-interlaced) on top of zlib, and JPEG (baseline / progressive, jpeg_reader.h:
// pixel-identical to the decoder the reference uses, so its JPEG-based known answers hold here).
// All SSIM work goes through the public API of this repository (include/rmgr/ssim.h, include/ssim_cuda.h);
// the -y luma conversion runs on the GPU (ssim_cuda_compute_luma).
#include <rmgr/ssim.h>
#include <rmgr/ssim-openmp.h>
#include <ssim_cuda.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jpeg_reader.h"

namespace
{

struct Image
{
    int width = 0, height = 0, channels = 0;
    std::vector<uint8_t> pixels;      // interleaved, row-major, top-down
};

std::string g_error;

bool read_file(const char* path, std::vector<uint8_t>& data)
{
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Failed to open file \"%s\"\n", path); return false; }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(size > 0 ? (size_t)size : 0);
    const bool ok = size >= 0 && fread(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok) fprintf(stderr, "Failed to read file \"%s\"\n", path);
    return ok;
}

// ---------------------------------------------------------------------------------------------- PNM
bool pnm_token(const std::vector<uint8_t>& d, size_t& pos, int& value)
{
    for (;;) {
        while (pos < d.size() && isspace(d[pos])) ++pos;
        if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') ++pos; continue; }
        break;
    }
    if (pos >= d.size() || !isdigit(d[pos])) return false;
    long v = 0;
    while (pos < d.size() && isdigit(d[pos])) { v = v * 10 + (d[pos++] - '0'); if (v > 1 << 30) return false; }
    value = (int)v;
    return true;
}

bool decode_pnm(const std::vector<uint8_t>& d, Image& img)
{
    img.channels = (d[1] == '5') ? 1 : 3;
    size_t pos = 2;
    int maxval = 0;
    if (!pnm_token(d, pos, img.width) || !pnm_token(d, pos, img.height) || !pnm_token(d, pos, maxval)) { g_error = "bad PNM header"; return false; }
    if (maxval != 255) { g_error = "only 8-bit PNM (maxval 255) is supported"; return false; }
    ++pos;                                                           // single whitespace after maxval
    const size_t need = (size_t)img.width * img.height * img.channels;
    if (img.width <= 0 || img.height <= 0 || d.size() < pos + need) { g_error = "truncated PNM"; return false; }
    img.pixels.assign(d.begin() + pos, d.begin() + pos + need);
    return true;
}

// ---------------------------------------------------------------------------------------------- PNG
uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

bool decode_png(const std::vector<uint8_t>& d, Image& img)
{
    size_t pos = 8;
    int bitDepth = 0, colorType = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    bool haveHeader = false;
    while (pos + 12 <= d.size()) {
        const uint32_t len = be32(&d[pos]);
        const uint8_t* type = &d[pos + 4];
        const uint8_t* body = &d[pos + 8];
        if (pos + 12 + (size_t)len > d.size()) { g_error = "truncated PNG chunk"; return false; }
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            img.width = (int)be32(body); img.height = (int)be32(body + 4);
            bitDepth = body[8]; colorType = body[9]; interlace = body[12];
            haveHeader = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!haveHeader || img.width <= 0 || img.height <= 0) { g_error = "bad PNG header"; return false; }
    if (bitDepth != 8 || interlace != 0) { g_error = "only 8-bit non-interlaced PNG is supported"; return false; }
    int srcChannels;
    switch (colorType) {
        case 0: srcChannels = 1; break;
        case 2: srcChannels = 3; break;
        case 3: srcChannels = 1; break;
        case 4: srcChannels = 2; break;
        case 6: srcChannels = 4; break;
        default: g_error = "unsupported PNG colour type"; return false;
    }
    const size_t rowBytes = (size_t)img.width * srcChannels;
    std::vector<uint8_t> raw((rowBytes + 1) * img.height);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) { g_error = "PNG inflate failed"; return false; }

    // undo the per-scanline filters (PNG specification, section 9)
    std::vector<uint8_t> out(rowBytes * img.height);
    const int bpp = srcChannels;
    for (int y = 0; y < img.height; ++y) {
        const uint8_t filter = raw[(rowBytes + 1) * y];
        const uint8_t* src = &raw[(rowBytes + 1) * y + 1];
        uint8_t* dst = &out[rowBytes * y];
        const uint8_t* up = y ? dst - rowBytes : nullptr;
        for (size_t x = 0; x < rowBytes; ++x) {
            const int a = x >= (size_t)bpp ? dst[x - bpp] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                default: g_error = "bad PNG filter"; return false;
            }
            dst[x] = (uint8_t)(src[x] + pred);
        }
    }
    if (colorType == 3) {                                            // palette -> RGB
        img.channels = 3;
        img.pixels.resize((size_t)img.width * img.height * 3);
        for (size_t i = 0; i < out.size(); ++i) {
            const size_t e = (size_t)out[i] * 3;
            for (int k = 0; k < 3; ++k) img.pixels[i * 3 + k] = e + k < palette.size() ? palette[e + k] : 0;
        }
    } else {
        img.channels = srcChannels;
        img.pixels.swap(out);
    }
    return true;
}

bool load_img(const char* path, Image& img)
{
    std::vector<uint8_t> d;
    if (!read_file(path, d)) return false;
    static const uint8_t pngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok;
    if (d.size() > 8 && !memcmp(d.data(), pngSig, 8)) ok = decode_png(d, img);
    else if (d.size() > 2 && d[0] == 'P' && (d[1] == '5' || d[1] == '6')) ok = decode_pnm(d, img);
    else if (d.size() > 3 && d[0] == 0xFF && d[1] == 0xD8 && d[2] == 0xFF) {
        jpegr::Decoder dec;
        ok = dec.decode(d.data(), d.size());
        if (ok) { img.width = dec.width; img.height = dec.height; img.channels = dec.channels; img.pixels.swap(dec.pixels); }
        else g_error = dec.error;
    }
    else { g_error = "unknown image format (supported: PNG, JPEG, binary PGM/PPM)"; ok = false; }
    if (!ok) fprintf(stderr, "Failed to load image \"%s\":\n%s\n", path, g_error.c_str());
    return ok;
}

// ---------------------------------------------------------------------------------------------- writers
void put_be32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }

void png_chunk(FILE* f, const char* type, const std::vector<uint8_t>& body)
{
    std::vector<uint8_t> head;
    put_be32(head, (uint32_t)body.size());
    fwrite(head.data(), 1, 4, f);
    uLong crc = crc32(0L, (const Bytef*)type, 4);
    if (!body.empty()) crc = crc32(crc, body.data(), (uInt)body.size());
    fwrite(type, 1, 4, f);
    if (!body.empty()) fwrite(body.data(), 1, body.size(), f);
    std::vector<uint8_t> tail;
    put_be32(tail, (uint32_t)crc);
    fwrite(tail.data(), 1, 4, f);
}

bool write_png(FILE* f, int w, int h, int c, const uint8_t* px)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    static const uint8_t colorTypes[5] = {0, 0, 4, 2, 6};
    fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(colorTypes[c]); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(f, "IHDR", ihdr);
    const size_t rowBytes = (size_t)w * c;
    std::vector<uint8_t> raw((rowBytes + 1) * h);
    for (int y = 0; y < h; ++y) { raw[(rowBytes + 1) * y] = 0; memcpy(&raw[(rowBytes + 1) * y + 1], px + rowBytes * y, rowBytes); }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    z.resize(zlen);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", std::vector<uint8_t>());
    return true;
}

void put_le16(FILE* f, unsigned v) { fputc(v & 255, f); fputc((v >> 8) & 255, f); }
void put_le32(FILE* f, unsigned v) { put_le16(f, v & 0xffff); put_le16(f, v >> 16); }

void write_tga(FILE* f, int w, int h, int c, const uint8_t* px)
{
    // uncompressed, top-left origin; 1 channel = grayscale (type 3), 3/4 = true colour (type 2, BGR(A))
    const uint8_t head[12] = {0, 0, (uint8_t)(c <= 2 ? 3 : 2), 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(head, 1, 12, f);
    put_le16(f, w); put_le16(f, h);
    fputc(8 * c, f);
    fputc(0x20 | (c == 2 || c == 4 ? 8 : 0), f);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const uint8_t* p = px + i * c;
        if (c <= 2) fwrite(p, 1, c, f);
        else { fputc(p[2], f); fputc(p[1], f); fputc(p[0], f); if (c == 4) fputc(p[3], f); }
    }
}

void write_bmp(FILE* f, int w, int h, int c, const uint8_t* px)
{
    // 24-bit BGR, bottom-up, rows padded to 4 bytes (gray is replicated; alpha is dropped)
    const unsigned rowBytes = ((unsigned)w * 3 + 3) & ~3u, size = 54 + rowBytes * h;
    fputc('B', f); fputc('M', f); put_le32(f, size); put_le32(f, 0); put_le32(f, 54);
    put_le32(f, 40); put_le32(f, w); put_le32(f, h); put_le16(f, 1); put_le16(f, 24);
    put_le32(f, 0); put_le32(f, rowBytes * h); put_le32(f, 2835); put_le32(f, 2835); put_le32(f, 0); put_le32(f, 0);
    std::vector<uint8_t> row(rowBytes, 0);
    for (int y = h - 1; y >= 0; --y) {
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = px + ((size_t)y * w + x) * c;
            const uint8_t r = p[0], g = c >= 3 ? p[1] : p[0], b = c >= 3 ? p[2] : p[0];
            row[x * 3 + 0] = b; row[x * 3 + 1] = g; row[x * 3 + 2] = r;
        }
        fwrite(row.data(), 1, rowBytes, f);
    }
}

void print_help(FILE* file)
{
    fprintf(file, "Usage: rmgr-ssim [options] img1 img2 [map]\n"
                  "Options:\n"
                  "  -#  Compute SSIM only for channel #\n"
                  "  -y  Compute SSIM on luminance\n"
                  "      For images with <= 2 channels, only channel 0's SSIM will be computed\n"
                  "      For images with >= 3 channels, first three channels are converted from RGB to Y\n\n"
                  "Images: PNG (8-bit, non-interlaced), JPEG or binary PGM/PPM.  Map: .pfm, .png, .bmp, .tga, .pgm/.ppm\n"
                  "Backend: ssim_b200 (CUDA sm_100a), device selected by SSIM_CUDA_DEVICE\n");
}

// one channel of an interleaved pair, same parameter set-up as the reference CLI (src/ssim-cli.cpp:108-127)
int32_t compute_channel(float* ssim, const Image& a, const Image& b, int channel, float* map, int mapChannels, int mapChannel)
{
    rmgr::ssim::GeneralParams params;
    memset(&params, 0, sizeof(params));
    params.width  = (uint32_t)a.width;
    params.height = (uint32_t)a.height;
    params.imgA.init_interleaved(a.pixels.data(), (ptrdiff_t)a.width * a.channels, a.channels, channel);
    params.imgB.init_interleaved(b.pixels.data(), (ptrdiff_t)b.width * b.channels, b.channels, channel);
    params.ssimMap    = map ? map + mapChannel : NULL;
    params.ssimStep   = mapChannels;
    params.ssimStride = (ptrdiff_t)a.width * mapChannels;
    return rmgr::ssim::compute_ssim_openmp(ssim, params);
}

int report(int32_t rc)
{
    fprintf(stderr, "SSIM computation failed: %s (%s)\n", strerror(rc), ssim_cuda_last_error_string());
    return EXIT_FAILURE;
}

} // namespace


int main(int argc, char* argv[])
{
    if (argc == 2 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))) { print_help(stdout); return EXIT_SUCCESS; }
    if (argc == 3 && !strcmp(argv[1], "--probe")) {
        // decoder self-check (no GPU needed): prints "width height channels fnv1a64"
        Image img;
        if (!load_img(argv[2], img)) return EXIT_FAILURE;
        uint64_t hsh = 1469598103934665603ull;
        for (uint8_t v : img.pixels) hsh = (hsh ^ v) * 1099511628211ull;
        printf("%d %d %d %016llx\n", img.width, img.height, img.channels, (unsigned long long)hsh);
        return EXIT_SUCCESS;
    }
    if (argc < 3 || argc > 5) { print_help(stderr); return EXIT_FAILURE; }

    int onlyChannel = -1, filesIndex = 1;
    bool luminance = false;
    if (argc >= 4 && argv[1][0] == '-') {
        const char* option = argv[1];
        if (option[1] >= '0' && option[1] <= '3' && option[2] == 0) onlyChannel = option[1] - '0';
        else if (!strcmp(option, "-y")) luminance = true;
        else { fprintf(stderr, "Unknown option: %s\n", option); return EXIT_FAILURE; }
        filesIndex = 2;
    }
    const char* mapPath = (argc - filesIndex == 3) ? argv[filesIndex + 2] : NULL;

    Image img1, img2;
    if (!load_img(argv[filesIndex], img1) || !load_img(argv[filesIndex + 1], img2)) return EXIT_FAILURE;
    if (img1.width != img2.width || img1.height != img2.height) {
        fprintf(stderr, "Images do not have the same dimensions: %ux%u vs %ux%u\n", img1.width, img1.height, img2.width, img2.height);
        return EXIT_FAILURE;
    }
    if (img1.channels != img2.channels) {
        fprintf(stderr, "Images do not have the same number of channels: %u vs %u\n", img1.channels, img2.channels);
        return EXIT_FAILURE;
    }
    if (onlyChannel >= img1.channels) {
        fprintf(stderr, "Cannot compute SSIM for channel %u, images have only %u channels\n", onlyChannel, img1.channels);
        return EXIT_FAILURE;
    }
    const int W = img1.width, H = img1.height, C = img1.channels;
    if (C < 3 && luminance) { onlyChannel = 0; luminance = false; }

    const int mapChannels = mapPath ? ((onlyChannel >= 0 || luminance) ? 1 : C) : 0;
    std::vector<float> map((size_t)W * H * mapChannels);
    float* mapPtr = mapPath ? map.data() : NULL;

    if (onlyChannel >= 0) {
        float ssim;
        const int32_t rc = compute_channel(&ssim, img1, img2, onlyChannel, mapPtr, mapChannels, 0);
        if (rc != 0) return report(rc);
        printf("% 7.4f\n", ssim);
    } else if (luminance) {
        float ssim;
        const char* env = getenv("SSIM_CUDA_DEVICE");
        const int32_t rc = ssim_cuda_compute_luma(env ? atoi(env) : 0, (uint32_t)W, (uint32_t)H, img1.pixels.data(), C, (ptrdiff_t)W * C,
                                                  img2.pixels.data(), C, (ptrdiff_t)W * C, mapPtr, 1, W, &ssim);
        if (rc != 0) return report(rc);
        printf("% 7.4f\n", ssim);
    } else {
        // every channel (src/ssim-cli.cpp:199-209 loops compute_ssim over the channels; here one pass does them all)
        float ssims[16], average = 0.0f;
        const char* env = getenv("SSIM_CUDA_DEVICE");
        const int32_t rc = ssim_cuda_compute_channels(env ? atoi(env) : 0, (uint32_t)W, (uint32_t)H, (uint32_t)C, img1.pixels.data(), (ptrdiff_t)W * C,
                                                      img2.pixels.data(), (ptrdiff_t)W * C, mapPtr, (ptrdiff_t)W * C, ssims);
        if (rc != 0) return report(rc);
        for (int c = 0; c < C; ++c) {
            printf("Channel %u: % 7.4f\n", c, ssims[c]);
            average += ssims[c];
        }
        printf("Average  : % 7.4f\n", average / C);
    }

    if (mapPath == NULL) return EXIT_SUCCESS;

    // ---- map output (reference src/ssim-cli.cpp:298-383): format from the extension, 8-bit = max(0, s) * 255
    const char* ext = strrchr(mapPath, '.');
    enum { F_TGA, F_BMP, F_PNG, F_PFM, F_PNM } fmt = F_TGA;
    if (ext == NULL) fprintf(stderr, "Cannot deduce file format from extension, saving as tga\n");
    else if (!strcasecmp(ext, ".bmp")) fmt = F_BMP;
    else if (!strcasecmp(ext, ".png")) fmt = F_PNG;
    else if (!strcasecmp(ext, ".tga")) fmt = F_TGA;
    else if (!strcasecmp(ext, ".pgm") || !strcasecmp(ext, ".ppm")) fmt = F_PNM;
    else if (!strcasecmp(ext, ".pfm")) fmt = F_PFM;
    else return EXIT_FAILURE;
    if ((fmt == F_PFM || fmt == F_PNM) && mapChannels != 1 && mapChannels != 3) {
        fprintf(stderr, "PFM/PNM images can only contain 1 or 3 channels but the map contains %d channels\n", mapChannels);
        return EXIT_FAILURE;
    }
    FILE* f = fopen(mapPath, "wb");
    if (f == NULL) { fprintf(stderr, "Failed to open file \"%s\" for writing\n", mapPath); return EXIT_FAILURE; }
    int retval = EXIT_SUCCESS;
    if (fmt == F_PFM) {
        fprintf(f, "P%c\n%d %d\n-1.0\n", mapChannels == 1 ? 'f' : 'F', W, H);       // little endian, bottom-up
        const size_t stride = (size_t)W * mapChannels;
        for (int y = H; --y >= 0;)
            if (fwrite(map.data() + y * stride, sizeof(float), stride, f) != stride) { fprintf(stderr, "Error writing to file \"%s\"\n", mapPath); retval = EXIT_FAILURE; break; }
    } else {
        std::vector<uint8_t> map8(map.size());
        for (size_t i = 0; i < map.size(); ++i) map8[i] = (uint8_t)(std::max(0.0f, map[i]) * 255.0f);
        switch (fmt) {
            case F_PNG: if (!write_png(f, W, H, mapChannels, map8.data())) retval = EXIT_FAILURE; break;
            case F_BMP: write_bmp(f, W, H, mapChannels, map8.data()); break;
            case F_TGA: write_tga(f, W, H, mapChannels, map8.data()); break;
            default:    fprintf(f, "P%c\n%d %d\n255\n", mapChannels == 1 ? '5' : '6', W, H); fwrite(map8.data(), 1, map8.size(), f); break;
        }
    }
    fclose(f);
    return retval;
}
