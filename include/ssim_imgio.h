/* ssim_imgio.h -- C ABI of the image readers the `rmgr-ssim` front end carries (ssim_b200/lib/libssim_imgio.so).
 *
 * The reference reads its inputs through stb_image (src/ssim-cli.cpp:33-40, 143-156; tests/rmgr-ssim-tests.cpp:237-243),
 * which is fetched at configure time and absent offline.  This library replaces that one call for JPEG files with a reader
 * whose pixels are identical to that decoder's (ssim_b200/csrc/jpeg_reader.h), so the reference's JPEG-based known answers
 * (tests/rmgr-ssim-tests.cpp:388-465) hold for this repository too.  Host code only; no GPU involved. */
#ifndef SSIM_IMGIO_H
#define SSIM_IMGIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Decodes a baseline or progressive Huffman JPEG (8 bits per sample, 1 or 3 components) held in memory.
 * Replaces stbi_load(path, &w, &h, &channels, 0) of the reference (src/ssim-cli.cpp:143).
 *   out == NULL : only *width, *height, *channels are filled (the file is parsed completely);
 *   out != NULL : must hold width*height*channels bytes (capacity given in out_capacity); receives the pixels interleaved
 *                 (gray, or RGB), rows top-down, no padding.
 * Returns 0, EINVAL (not a JPEG / corrupt / unsupported variant / more than 2^28 pixels: see ssim_imgio_last_error()),
 * ERANGE (out too small) or ENOMEM. */
int ssim_imgio_decode_jpeg(const uint8_t* data, size_t size, uint8_t* out, size_t out_capacity,
                           int* width, int* height, int* channels);

/* Message of the last failure on the calling thread ("" if none). */
const char* ssim_imgio_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* SSIM_IMGIO_H */
