// ref_prelude.hpp — TEST INFRASTRUCTURE ONLY: glue between unity_shim.hpp and the generated translation of the reference.
#pragma once
#include "unity_shim.hpp"
namespace cpuvox_ref {
template <class A, class B>
inline int cs_compare(A a, B b) { return a < b ? -1 : (a > b ? 1 : 0); }

// System.IO.MemoryMappedFiles / FileInfo as WorldSaveFile.cs uses them: a file mapped with mmap; the objects unmap / close in their
// destructors (`using (...) { }` becomes a scope)
}  // namespace cpuvox_ref
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
namespace cpuvox_ref {
enum class FileMode { CreateNew = 1, Create = 2, Open = 3 };
struct FileInfo {
    long Length = 0;
    explicit FileInfo(const string& path) {
        struct stat st;
        if (stat(path.c_str(), &st) != 0) throw cs_exception("FileNotFoundException");
        Length = (long)st.st_size;
    }
};
struct SafeMemoryMappedViewHandle_ {
    byte* base = nullptr;
    void AcquirePointer(byte*& p) { p = base; }
    void ReleasePointer() {}
};
struct MemoryMappedViewAccessor {
    SafeMemoryMappedViewHandle_ SafeMemoryMappedViewHandle;
};
struct MemoryMappedFile {
    int fd = -1;
    byte* base = nullptr;
    long size = 0;
    MemoryMappedFile() {}
    MemoryMappedFile(const MemoryMappedFile&) = delete;
    MemoryMappedFile(MemoryMappedFile&& o) : fd(o.fd), base(o.base), size(o.size) { o.fd = -1; o.base = nullptr; }
    static MemoryMappedFile CreateFromFile(const string& path, FileMode mode, std::nullptr_t, long capacity) {
        MemoryMappedFile f;
        f.fd = open(path.c_str(), mode == FileMode::Open ? O_RDWR : (O_RDWR | O_CREAT | O_TRUNC), 0644);
        if (f.fd < 0) throw cs_exception("IOException: cannot open file");
        if (mode != FileMode::Open && ftruncate(f.fd, capacity) != 0) throw cs_exception("IOException: ftruncate");
        f.size = capacity;
        void* p = mmap(nullptr, (size_t)capacity, PROT_READ | PROT_WRITE, MAP_SHARED, f.fd, 0);
        if (p == MAP_FAILED) throw cs_exception("IOException: mmap");
        f.base = (byte*)p;
        return f;
    }
    MemoryMappedViewAccessor CreateViewAccessor() { MemoryMappedViewAccessor a; a.SafeMemoryMappedViewHandle.base = base; return a; }
    ~MemoryMappedFile() {
        if (base) { msync(base, (size_t)size, MS_SYNC); munmap(base, (size_t)size); }
        if (fd >= 0) close(fd);
    }
};
}  // namespace cpuvox_ref

// ---------------------------------------------------------------------------------------------------------------------
// Presentation side of UnityEngine that RenderManager drives (Mesh / Material / CommandBuffer / Screen): host stand-ins.
// CommandBuffer.DrawMesh is a small software rasteriser (ours: a GPU's interpolator bits are not reproducible anyway) that
// interpolates the vertex attributes RenderManager.BlitSegments set up and calls the fragment function translated from
// the reference's RayBufferBlit.shader for every covered pixel.  D3D conventions: SV_POSITION is
// the pixel centre, y counted from the top; the render target here is stored bottom-up (row 0 = bottom, Unity screen space).
// ---------------------------------------------------------------------------------------------------------------------
namespace cpuvox_ref {

struct Screen {
    static inline int width = 0, height = 0;
};
struct Bounds {
    Bounds() {}
    Bounds(Vector3, Vector3) {}
};
enum class MeshTopology { Triangles };
enum class CameraEvent { AfterForwardOpaque };
struct Mesh {
    std::vector<float3> vertices;
    std::vector<float4> uv;
    std::vector<ushort> indices;
    Bounds bounds;
    void SetVertices(const NativeArray<float3>& v) { vertices.assign(v.ptr, v.ptr + v.Length); }
    void SetUVs(int, const NativeArray<float4>& v, int start, int count) { uv.assign(v.ptr + start, v.ptr + start + count); }
    void SetIndices(const NativeArray<ushort>& v, MeshTopology, int, bool, int) { indices.assign(v.ptr, v.ptr + v.Length); }
    void UploadMeshData(bool) {}
};
struct Material {
    RenderTexture tex1, tex2;
    Vector4 rayOffset, rayScale;
    void SetTexture(const char* name, const RenderTexture& t) { (strcmp(name, "_MainTex1") == 0 ? tex1 : tex2) = t; }
    void SetVector(const char* name, Vector4 v) { (strcmp(name, "_RayOffset") == 0 ? rayOffset : rayScale) = v; }
};

// what the translated fragment function sees
struct v2f {
    float4 vertex;  // SV_POSITION
    float4 uv;      // TEXCOORD0
};
struct ShaderGlobals {
    const RenderTexture* _MainTex1;
    const RenderTexture* _MainTex2;
    Vector4 _RayOffset, _RayScale;
    float4 _ScreenParams;
};
// tex2D on a point-filtered, clamped ARGB32 texture: channels as bytes / 255 in a,r,g,b order (kept exact by pack_argb)
inline float4 tex2D(const RenderTexture* t, float2 uv) {
    int x = cs_f2i(std::floor(uv.x * (float)t->width)), y = cs_f2i(std::floor(uv.y * (float)t->height));
    x = x < 0 ? 0 : (x >= t->width ? t->width - 1 : x);
    y = y < 0 ? 0 : (y >= t->height ? t->height - 1 : y);
    uint32_t p = (*t->px)[(size_t)y * t->width + x];
    return float4((float)(p & 255u), (float)((p >> 8) & 255u), (float)((p >> 16) & 255u), (float)(p >> 24));
}
inline float4 lerp(float4 a, float4 b, bool t) { return t ? b : a; }  // HLSL lerp(a, b, (float)cond) with cond in {0, 1}
inline uint32_t pack_argb(float4 c) { return (uint32_t)c.x | ((uint32_t)c.y << 8) | ((uint32_t)c.z << 16) | ((uint32_t)c.w << 24); }
float4 RayBufferBlit_frag(const v2f& i, const ShaderGlobals& g);  // generated from RayBufferBlit.shader

struct CommandBuffer {
    // render target of the camera the buffer is attached to: W x H, row 0 = bottom
    uint32_t* target = nullptr;
    int targetW = 0, targetH = 0;
    void Clear() {}
    void Dispose() {}
    void CopyTexture(const Texture2D& src, int, int, int srcX, int srcY, int w, int h, RenderTexture& dst, int, int, int dstX, int dstY) {
        for (int r = 0; r < h; r++)
            memcpy(dst.px->data() + (size_t)(dstY + r) * dst.width + dstX, src.pixels + (size_t)(srcY + r) * src.width + srcX, (size_t)w * 4);
    }
    // Rasterisation the way D3D11 specifies it: vertex positions snapped to 1/256 pixel, exact integer edge functions,
    // top-left fill rule (so the four segment triangles tile the screen without gaps or double hits), attributes
    // evaluated as a plane through the first vertex (a constant attribute such as the segment index stays exact).
    void DrawMesh(const Mesh& mesh, const Matrix4x4&, const Material& mat, int) {
        if (!target) return;
        const int W = targetW, H = targetH;
        ShaderGlobals g{&mat.tex1, &mat.tex2, mat.rayOffset, mat.rayScale, float4((float)W, (float)H, 1.0f + 1.0f / W, 1.0f + 1.0f / H)};
        for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
            int i0 = mesh.indices[t], i1 = mesh.indices[t + 1], i2 = mesh.indices[t + 2];
            auto snap = [](float v) -> int64_t {
                double s = std::nearbyint((double)v * 256.0);
                if (!(s > -4.0e12 && s < 4.0e12)) return INT64_MIN;
                return (int64_t)s;
            };
            // viewport transform; y counted from the top (SV_POSITION)
            int64_t X[3], Y[3];
            int idx[3] = {i0, i1, i2};
            bool bad = false;
            for (int k = 0; k < 3; k++) {
                const float3& p = mesh.vertices[idx[k]];
                X[k] = snap((p.x * 0.5f + 0.5f) * (float)W);
                Y[k] = snap((1.0f - (p.y * 0.5f + 0.5f)) * (float)H);
                bad |= X[k] == INT64_MIN || Y[k] == INT64_MIN;
            }
            if (bad) continue;
            __int128 area = (__int128)(X[1] - X[0]) * (Y[2] - Y[0]) - (__int128)(X[2] - X[0]) * (Y[1] - Y[0]);
            if (area == 0) continue;  // degenerate: an unused segment (all three vertices coincide or are collinear)
            if (area < 0) { std::swap(X[1], X[2]); std::swap(Y[1], Y[2]); std::swap(idx[1], idx[2]); area = -area; }
            const float4 a0 = mesh.uv[idx[0]], d1 = mesh.uv[idx[1]] - a0, d2 = mesh.uv[idx[2]] - a0;
            auto top_left = [](int64_t dx, int64_t dy) { return (dy == 0 && dx > 0) || dy < 0; };
            const bool tl01 = top_left(X[1] - X[0], Y[1] - Y[0]), tl12 = top_left(X[2] - X[1], Y[2] - Y[1]), tl20 = top_left(X[0] - X[2], Y[0] - Y[2]);
            const double inv_area = 1.0 / (double)area;
            auto rows = [&](int yb, int ye) {
                for (int yd = yb; yd < ye; yd++) {
                    const int64_t py = (int64_t)yd * 256 + 128;
                    for (int x = 0; x < W; x++) {
                        const int64_t px = (int64_t)x * 256 + 128;
                        __int128 e01 = (__int128)(X[1] - X[0]) * (py - Y[0]) - (__int128)(Y[1] - Y[0]) * (px - X[0]);
                        __int128 e12 = (__int128)(X[2] - X[1]) * (py - Y[1]) - (__int128)(Y[2] - Y[1]) * (px - X[1]);
                        __int128 e20 = (__int128)(X[0] - X[2]) * (py - Y[2]) - (__int128)(Y[0] - Y[2]) * (px - X[2]);
                        if (e01 < 0 || e12 < 0 || e20 < 0) continue;
                        if ((e01 == 0 && !tl01) || (e12 == 0 && !tl12) || (e20 == 0 && !tl20)) continue;
                        float w1 = (float)((double)e20 * inv_area), w2 = (float)((double)e01 * inv_area);
                        v2f in;
                        in.vertex = float4((float)x + 0.5f, (float)yd + 0.5f, 0.5f, 1.0f);
                        in.uv = a0 + d1 * float4(w1) + d2 * float4(w2);
                        target[(size_t)(H - 1 - yd) * W + x] = pack_argb(RayBufferBlit_frag(in, g));
                    }
                }
            };
            int threads = std::max(1, std::min(g_job_threads, H / 8));
            if (threads <= 1) rows(0, H);
            else {
                std::vector<std::thread> pool;
                for (int th = 0; th < threads; th++) pool.emplace_back(rows, (int)((int64_t)H * th / threads), (int)((int64_t)H * (th + 1) / threads));
                for (auto& th : pool) th.join();
            }
        }
    }
};
}  // namespace cpuvox_ref
