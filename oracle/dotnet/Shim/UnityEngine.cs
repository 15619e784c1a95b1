// TEST INFRASTRUCTURE ONLY — stand-in for the UnityEngine types the linked reference files touch. Closed source: behaviour as
// documented (SURVEY.md Appendix A2-A11), same content as the C++ stand-in (oracle/refbuild/unity_shim.hpp, ref_prelude.hpp).
// Textures are host memory; CommandBuffer.CopyTexture copies at once, DrawMesh is not rasterised here (the frame comes from
// oracle/_ref). Never compiled in this image (no C# toolchain).
using System;
using System.Collections.Generic;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Mathematics;

namespace UnityEngine
{
    public static class Mathf
    {
        public const float Deg2Rad = (float)Math.PI * 2F / 360F;
        public const float Rad2Deg = 1F / Deg2Rad;
        public static int RoundToInt(float f) => (int)Math.Round(f);
        public static float Sin(float f) => (float)Math.Sin(f);
        public static float Cos(float f) => (float)Math.Cos(f);
        public static float Tan(float f) => (float)Math.Tan(f);
        public static float Acos(float f) => (float)Math.Acos(f);
        public static float Sqrt(float f) => (float)Math.Sqrt(f);
        public static float Abs(float f) => Math.Abs(f);
        public static float Sign(float f) => f >= 0F ? 1F : -1F;
        public static int Max(int a, int b) => a > b ? a : b;
        public static int Min(int a, int b) => a < b ? a : b;
        public static float Max(float a, float b) => a > b ? a : b;
        public static float Min(float a, float b) => a < b ? a : b;
        public static float Clamp(float v, float a, float b) => v < a ? a : (v > b ? b : v);
        public static float Clamp01(float v) => v < 0F ? 0F : (v > 1F ? 1F : v);
        public static int NextPowerOfTwo(int v) { v -= 1; v |= v >> 16; v |= v >> 8; v |= v >> 4; v |= v >> 2; v |= v >> 1; return v + 1; }
    }

    public struct Vector2
    {
        public float x, y;
        public Vector2(float x, float y) { this.x = x; this.y = y; }
        public static float Angle(Vector2 from, Vector2 to)
        {
            float denominator = (float)Math.Sqrt((from.x * from.x + from.y * from.y) * (to.x * to.x + to.y * to.y));
            if (denominator < 1e-15F) return 0F;
            float dot = Mathf.Clamp((from.x * to.x + from.y * to.y) / denominator, -1F, 1F);
            return (float)Math.Acos(dot) * Mathf.Rad2Deg;
        }
        public static float SignedAngle(Vector2 from, Vector2 to) => Angle(from, to) * Mathf.Sign(from.x * to.y - from.y * to.x);
    }

    public struct Vector3
    {
        public float x, y, z;
        public Vector3(float x, float y, float z) { this.x = x; this.y = y; this.z = z; }
        public static Vector3 zero => new Vector3(0, 0, 0);
        public static Vector3 one => new Vector3(1, 1, 1);
        public static Vector3 up => new Vector3(0, 1, 0);
        public static Vector3 forward => new Vector3(0, 0, 1);
        public static Vector3 operator +(Vector3 a, Vector3 b) => new Vector3(a.x + b.x, a.y + b.y, a.z + b.z);
        public static Vector3 operator -(Vector3 a, Vector3 b) => new Vector3(a.x - b.x, a.y - b.y, a.z - b.z);
        public static Vector3 operator *(Vector3 a, float d) => new Vector3(a.x * d, a.y * d, a.z * d);
        public static Vector3 operator *(float d, Vector3 a) => new Vector3(a.x * d, a.y * d, a.z * d);
        public static float Dot(Vector3 a, Vector3 b) => a.x * b.x + a.y * b.y + a.z * b.z;
        public static Vector3 Cross(Vector3 a, Vector3 b) => new Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
        public static float Magnitude(Vector3 a) => (float)Math.Sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
        public static Vector3 Normalize(Vector3 a) { float m = Magnitude(a); return m > 1E-05F ? new Vector3(a.x / m, a.y / m, a.z / m) : zero; }
        public static float Distance(Vector3 a, Vector3 b) { float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z; return (float)Math.Sqrt(dx * dx + dy * dy + dz * dz); }
    }

    public struct Vector4
    {
        public float x, y, z, w;
        public Vector4(float x, float y, float z, float w) { this.x = x; this.y = y; this.z = z; this.w = w; }
    }

    public struct Color32
    {
        public byte r, g, b, a;
        public Color32(byte r, byte g, byte b, byte a) { this.r = r; this.g = g; this.b = b; this.a = a; }
        public static implicit operator Color32(Color c) => new Color32(
            (byte)Math.Round(Mathf.Clamp01(c.r) * 255f), (byte)Math.Round(Mathf.Clamp01(c.g) * 255f),
            (byte)Math.Round(Mathf.Clamp01(c.b) * 255f), (byte)Math.Round(Mathf.Clamp01(c.a) * 255f));
        public static implicit operator Color(Color32 c) => new Color(c.r / 255f, c.g / 255f, c.b / 255f, c.a / 255f);
    }

    public struct Color
    {
        public float r, g, b, a;
        public Color(float r, float g, float b, float a = 1f) { this.r = r; this.g = g; this.b = b; this.a = a; }
        public static Color white => new Color(1, 1, 1, 1);
        public static Color red => new Color(1, 0, 0, 1);
        public static Color operator *(Color a, Color b) => new Color(a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a);
    }

    public struct Quaternion
    {
        public float x, y, z, w;
        public Quaternion(float x, float y, float z, float w) { this.x = x; this.y = y; this.z = z; this.w = w; }
        public static Quaternion identity => new Quaternion(0, 0, 0, 1);
        public static Vector3 operator *(Quaternion q, Vector3 p)
        {
            float x2 = q.x * 2F, y2 = q.y * 2F, z2 = q.z * 2F;
            float xx = q.x * x2, yy = q.y * y2, zz = q.z * z2, xy = q.x * y2, xz = q.x * z2, yz = q.y * z2, wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
            return new Vector3((1F - (yy + zz)) * p.x + (xy - wz) * p.y + (xz + wy) * p.z,
                               (xy + wz) * p.x + (1F - (xx + zz)) * p.y + (yz - wx) * p.z,
                               (xz - wy) * p.x + (yz + wx) * p.y + (1F - (xx + yy)) * p.z);
        }
    }

    public struct Matrix4x4
    {
        public float m00, m10, m20, m30, m01, m11, m21, m31, m02, m12, m22, m32, m03, m13, m23, m33;
        public static Matrix4x4 identity { get { Matrix4x4 m = default; m.m00 = m.m11 = m.m22 = m.m33 = 1F; return m; } }
        public static Matrix4x4 Scale(Vector3 s) { Matrix4x4 m = identity; m.m00 = s.x; m.m11 = s.y; m.m22 = s.z; return m; }
        public static Matrix4x4 LookAt(Vector3 from, Vector3 to, Vector3 up)
        {
            Vector3 f = Vector3.Normalize(to - from);
            Vector3 r = Vector3.Normalize(Vector3.Cross(up, f));
            Vector3 u = Vector3.Cross(f, r);
            Matrix4x4 m = identity;
            m.m00 = r.x; m.m10 = r.y; m.m20 = r.z;
            m.m01 = u.x; m.m11 = u.y; m.m21 = u.z;
            m.m02 = f.x; m.m12 = f.y; m.m22 = f.z;
            m.m03 = from.x; m.m13 = from.y; m.m23 = from.z;
            return m;
        }
    }

    public class Object { public static void Destroy(object o) { } }

    public class Transform
    {
        public Vector3 position;
        Quaternion rot = Quaternion.identity;
        public Quaternion rotation { get => rot; set => rot = value; }
        public Vector3 forward => rot * new Vector3(0, 0, 1);
        public Vector3 up => rot * new Vector3(0, 1, 0);
        public Vector3 eulerAngles
        {
            get
            {   // only .x is read (RenderManager.cs:377): Z-X-Y Euler pitch, sin(eulerAngles.x) = -forward.y
                float fy = Mathf.Clamp(forward.y, -1F, 1F);
                float pitch = (float)(Math.Asin(-fy) * (180.0 / Math.PI));
                if (pitch < 0F) pitch += 360F;
                return new Vector3(pitch, 0, 0);
            }
        }
    }

    public enum CameraEvent { AfterForwardOpaque }

    public class Camera
    {
        public Transform transform = new Transform();
        public float nearClipPlane = 0.05f, farClipPlane = 1000f, fieldOfView = 60f;
        public int pixelWidth, pixelHeight;
        public Matrix4x4 nonJitteredProjectionMatrix
        {
            get
            {
                float aspect = (float)pixelWidth / pixelHeight;
                float cot = 1F / (float)Math.Tan(fieldOfView * Mathf.Deg2Rad * 0.5F);
                Matrix4x4 p = default;
                p.m00 = cot / aspect; p.m11 = cot;
                p.m22 = -(farClipPlane + nearClipPlane) / (farClipPlane - nearClipPlane);
                p.m23 = -(2F * farClipPlane * nearClipPlane) / (farClipPlane - nearClipPlane);
                p.m32 = -1F;
                return p;
            }
        }
        public Matrix4x4 worldToCameraMatrix
        {
            get
            {
                Quaternion q = transform.rotation;
                Vector3 r = q * new Vector3(1, 0, 0), u = q * new Vector3(0, 1, 0), f = q * new Vector3(0, 0, 1), pos = transform.position;
                Matrix4x4 v = Matrix4x4.identity;
                v.m00 = r.x; v.m01 = r.y; v.m02 = r.z; v.m03 = -Vector3.Dot(r, pos);
                v.m10 = u.x; v.m11 = u.y; v.m12 = u.z; v.m13 = -Vector3.Dot(u, pos);
                v.m20 = -f.x; v.m21 = -f.y; v.m22 = -f.z; v.m23 = Vector3.Dot(f, pos);
                return v;
            }
        }
        public void RemoveAllCommandBuffers() { }
        public void AddCommandBuffer(CameraEvent e, Rendering.CommandBuffer b) { }
    }

    public static class Screen { public static int width, height; }
    public static class Debug
    {
        public static void DrawLine(Vector3 a, Vector3 b) { }
        public static void DrawLine(Vector3 a, Vector3 b, Color c) { }
        public static void DrawLine(Vector2 a, Vector2 b) { }
        public static void DrawLine(Vector2 a, Vector2 b, Color c) { }
        public static void Log(object o) { }
        public static void LogException(Exception e) { Console.Error.WriteLine(e); }
    }
    public struct Bounds { public Bounds(Vector3 c, Vector3 s) { } }
    public enum MeshTopology { Triangles }
    public enum FilterMode { Point, Bilinear }
    public enum TextureFormat { ARGB32 }
    public enum RenderTextureFormat { ARGB32 }
    public struct RenderTextureDescriptor
    {
        public int width, height;
        public RenderTextureDescriptor(int w, int h, RenderTextureFormat f, int depth, int mips) { width = w; height = h; }
    }

    public unsafe class Texture2D
    {
        public int width, height;
        public FilterMode filterMode;
        internal uint* pixels;
        public Texture2D(int w, int h) : this(w, h, TextureFormat.ARGB32, false, false) { }
        public Texture2D(int w, int h, TextureFormat f, bool mip, bool linear)
        {
            width = w; height = h;
            pixels = (uint*)UnsafeUtility.Malloc((long)w * h * 4, 16, Allocator.Persistent);
            UnsafeUtility.MemClear(pixels, (long)w * h * 4);
        }
        public NativeArray<T> GetRawTextureData<T>() where T : struct => NativeArrayUnsafeUtility.ConvertExistingDataToNativeArray<T>(pixels, width * height * 4 / UnsafeUtility.SizeOf<T>(), Allocator.None);
        public void LoadRawTextureData<T>(NativeArray<T> data) where T : struct => UnsafeUtility.MemCpy(pixels, data.GetUnsafePtr(), (long)width * height * 4);
        public void Apply(bool a, bool b) { }
        public Color32[] GetPixels32() { var r = new Color32[width * height]; return r; }
        public bool LoadImage(byte[] data, bool nonReadable) => false;
        ~Texture2D() { UnsafeUtility.Free(pixels, Allocator.Persistent); }
    }

    public unsafe class RenderTexture
    {
        public int width, height;
        public FilterMode filterMode;
        public uint* pixels;
        public RenderTexture(RenderTextureDescriptor d)
        {
            width = d.width; height = d.height;
            pixels = (uint*)UnsafeUtility.Malloc((long)width * height * 4, 16, Allocator.Persistent);
            UnsafeUtility.MemClear(pixels, (long)width * height * 4);
        }
        ~RenderTexture() { UnsafeUtility.Free(pixels, Allocator.Persistent); }
    }

    public class Mesh
    {
        public Bounds bounds;
        public float3[] vertices; public float4[] uv; public ushort[] indices;
        public void SetVertices(NativeArray<float3> v) { vertices = new float3[v.Length]; for (int i = 0; i < v.Length; i++) vertices[i] = v[i]; }
        public void SetUVs(int ch, NativeArray<float4> v, int start, int count) { uv = new float4[count]; for (int i = 0; i < count; i++) uv[i] = v[start + i]; }
        public void SetIndices(NativeArray<ushort> v, MeshTopology t, int sub, bool calcBounds, int baseVertex) { indices = new ushort[v.Length]; for (int i = 0; i < v.Length; i++) indices[i] = v[i]; }
        public void UploadMeshData(bool markNoLongerReadable) { }
    }

    public class Material
    {
        public Dictionary<string, RenderTexture> textures = new Dictionary<string, RenderTexture>();
        public Dictionary<string, Vector4> vectors = new Dictionary<string, Vector4>();
        public void SetTexture(string n, RenderTexture t) => textures[n] = t;
        public void SetVector(string n, Vector4 v) => vectors[n] = v;
    }

    public static class Graphics { public static void ExecuteCommandBuffer(Rendering.CommandBuffer b) { } }
}

namespace UnityEngine.Profiling
{
    public static class Profiler { public static void BeginSample(string s) { } public static void EndSample() { } }
}

namespace UnityEngine.Rendering
{
    public unsafe class CommandBuffer : IDisposable
    {
        public void Clear() { }
        public void Dispose() { }
        public void SetRenderTarget(RenderTexture t) { }
        public void ClearRenderTarget(bool depth, bool color, Color c) { }
        public void DrawMesh(Mesh m, Matrix4x4 mat, Material material, int submesh) { }
        public void CopyTexture(Texture2D src, int srcElement, int srcMip, int srcX, int srcY, int w, int h, RenderTexture dst, int dstElement, int dstMip, int dstX, int dstY)
        {
            for (int r = 0; r < h; r++)
                UnsafeUtility.MemCpy(dst.pixels + (long)(dstY + r) * dst.width + dstX, src.pixels + (long)(srcY + r) * src.width + srcX, (long)w * 4);
        }
    }
}

// the two members of UnityManager (a MonoBehaviour, not linked) that the linked files name
public class UnityManager
{
    public const int LOD_LEVELS = 6;   // UnityManager.cs:42
    public enum ERenderMode { ScreenBuffer, RayBufferTopDown, RayBufferLeftRight }
}
