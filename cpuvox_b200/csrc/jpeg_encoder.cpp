#include "jpeg_encoder.h"

#include <dlfcn.h>
#include <nvjpeg.h>

namespace cvxjpeg {

namespace {
// the entry points used, resolved from libnvjpeg.so.12 at run time
struct Api {
    void* lib = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegEncoderStateCreate) StateCreate = nullptr;
    decltype(&nvjpegEncoderStateDestroy) StateDestroy = nullptr;
    decltype(&nvjpegEncoderParamsCreate) ParamsCreate = nullptr;
    decltype(&nvjpegEncoderParamsDestroy) ParamsDestroy = nullptr;
    decltype(&nvjpegEncoderParamsSetQuality) SetQuality = nullptr;
    decltype(&nvjpegEncoderParamsSetSamplingFactors) SetSampling = nullptr;
    decltype(&nvjpegEncoderParamsSetOptimizedHuffman) SetOptimizedHuffman = nullptr;
    decltype(&nvjpegEncodeImage) EncodeImage = nullptr;
    decltype(&nvjpegEncodeRetrieveBitstream) RetrieveBitstream = nullptr;
};

template <typename F>
bool sym(void* lib, const char* name, F& f, std::string& error) {
    f = reinterpret_cast<F>(dlsym(lib, name));
    if (!f) { error = std::string("libnvjpeg: missing symbol ") + name; return false; }
    return true;
}

bool load(Api& a, std::string& error) {
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (a.lib) break;
    }
    if (!a.lib) { error = std::string("libnvjpeg.so.12 cannot be loaded: ") + dlerror(); return false; }
    return sym(a.lib, "nvjpegCreateSimple", a.CreateSimple, error) && sym(a.lib, "nvjpegDestroy", a.Destroy, error) &&
           sym(a.lib, "nvjpegEncoderStateCreate", a.StateCreate, error) && sym(a.lib, "nvjpegEncoderStateDestroy", a.StateDestroy, error) &&
           sym(a.lib, "nvjpegEncoderParamsCreate", a.ParamsCreate, error) && sym(a.lib, "nvjpegEncoderParamsDestroy", a.ParamsDestroy, error) &&
           sym(a.lib, "nvjpegEncoderParamsSetQuality", a.SetQuality, error) &&
           sym(a.lib, "nvjpegEncoderParamsSetSamplingFactors", a.SetSampling, error) &&
           sym(a.lib, "nvjpegEncoderParamsSetOptimizedHuffman", a.SetOptimizedHuffman, error) &&
           sym(a.lib, "nvjpegEncodeImage", a.EncodeImage, error) && sym(a.lib, "nvjpegEncodeRetrieveBitstream", a.RetrieveBitstream, error);
}

std::string status_text(const char* what, nvjpegStatus_t st) { return std::string(what) + " failed with nvjpegStatus " + std::to_string((int)st); }
} // namespace

struct Encoder {
    Api api;
    nvjpegHandle_t handle = nullptr;
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
};

void destroy(Encoder* e) {
    if (!e) return;
    if (e->params) e->api.ParamsDestroy(e->params);
    if (e->state) e->api.StateDestroy(e->state);
    if (e->handle) e->api.Destroy(e->handle);
    if (e->api.lib) dlclose(e->api.lib);
    delete e;
}

Encoder* create(std::string& error) {
    Encoder* e = new Encoder();
    nvjpegStatus_t st;
    if (!load(e->api, error)) { destroy(e); return nullptr; }
    if ((st = e->api.CreateSimple(&e->handle)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegCreateSimple", st); e->handle = nullptr; destroy(e); return nullptr; }
    if ((st = e->api.StateCreate(e->handle, &e->state, nullptr)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncoderStateCreate", st); e->state = nullptr; destroy(e); return nullptr; }
    if ((st = e->api.ParamsCreate(e->handle, &e->params, nullptr)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncoderParamsCreate", st); e->params = nullptr; destroy(e); return nullptr; }
    return e;
}

bool encode(Encoder* e, const uint8_t* rgb, int width, int height, int quality, int subsampling, cudaStream_t stream,
            std::vector<uint8_t>& out, std::string& error) {
    nvjpegStatus_t st;
    if ((st = e->api.SetQuality(e->params, quality, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncoderParamsSetQuality", st); return false; }
    if ((st = e->api.SetSampling(e->params, subsampling ? NVJPEG_CSS_420 : NVJPEG_CSS_444, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncoderParamsSetSamplingFactors", st); return false; }
    if ((st = e->api.SetOptimizedHuffman(e->params, 0, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncoderParamsSetOptimizedHuffman", st); return false; }
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; c++) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = const_cast<unsigned char*>(rgb);
    img.pitch[0] = (size_t)width * 3;
    if ((st = e->api.EncodeImage(e->handle, e->state, e->params, &img, NVJPEG_INPUT_RGBI, width, height, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncodeImage", st); return false; }
    size_t length = 0;
    if ((st = e->api.RetrieveBitstream(e->handle, e->state, nullptr, &length, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncodeRetrieveBitstream (size)", st); return false; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "cudaStreamSynchronize after nvjpegEncodeImage failed"; return false; }
    out.resize(length);
    if ((st = e->api.RetrieveBitstream(e->handle, e->state, out.data(), &length, stream)) != NVJPEG_STATUS_SUCCESS) { error = status_text("nvjpegEncodeRetrieveBitstream", st); return false; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { error = "cudaStreamSynchronize after nvjpegEncodeRetrieveBitstream failed"; return false; }
    out.resize(length);
    return true;
}

} // namespace cvxjpeg
