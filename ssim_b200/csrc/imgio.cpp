// imgio.cpp -- C ABI around the front end's JPEG reader (include/ssim_imgio.h), so that tests and bindings can check the
// decoded pixels against the reference's JPEG-based known answers without going through the command line.
#include <ssim_imgio.h>

#include <cerrno>
#include <new>

#include "jpeg_reader.h"

namespace { thread_local std::string t_error; }

extern "C" int ssim_imgio_decode_jpeg(const uint8_t* data, size_t size, uint8_t* out, size_t out_capacity,
                                      int* width, int* height, int* channels)
{
    t_error.clear();
    if (!data || !width || !height || !channels) { t_error = "null argument"; return EINVAL; }
    try {
        jpegr::Decoder dec;
        if (!dec.decode(data, size)) { t_error = dec.error; return EINVAL; }
        *width = dec.width; *height = dec.height; *channels = dec.channels;
        if (out) {
            if (out_capacity < dec.pixels.size()) { t_error = "output buffer too small"; return ERANGE; }
            memcpy(out, dec.pixels.data(), dec.pixels.size());
        }
        return 0;
    } catch (const std::bad_alloc&) { t_error = "out of memory"; return ENOMEM; }
}

extern "C" const char* ssim_imgio_last_error(void) { return t_error.c_str(); }
