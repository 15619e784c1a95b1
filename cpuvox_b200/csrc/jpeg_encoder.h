/* Presentation as a compressed still: the device frame -> baseline JPEG through nvJPEG's CUDA encoder ("encode of the
 * device framebuffer"; the step after RenderManager.cs:192-193, where the reference hands its target to Unity). libnvjpeg is a CUDA toolkit
 * library, loaded with dlopen on first use: the renderer itself has no link-time dependency on it, and a machine without it gets
 * CVX_ERR_UNSUPPORTED from cvx_present_jpeg and nothing else changes. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace cvxjpeg {

struct Encoder;                       // one nvJPEG handle + encoder state + parameter set (not thread safe: one per context)
Encoder* create(std::string& error);  // nullptr + error text when libnvjpeg.so.12 cannot be loaded or initialised
void destroy(Encoder* e);
/* rgb: device pointer to width*height interleaved R,G,B bytes, rows top-down, pitch width*3; ordered on `stream`. The bitstream lands in
 * `out` (host) and the stream has been synchronised when the call returns. subsampling: 0 = 4:4:4, 1 = 4:2:0. */
bool encode(Encoder* e, const uint8_t* rgb, int width, int height, int quality, int subsampling, cudaStream_t stream,
            std::vector<uint8_t>& out, std::string& error);

} // namespace cvxjpeg
