// TEST INFRASTRUCTURE ONLY — stand-in for Unity.Mathematics 1.2.6 (the subset the linked reference files call).
// Our restatement of the package's published scalar definitions (SURVEY.md Appendix A1); same content as the C++ stand-in
// oracle/refbuild/unity_shim.hpp, which IS compiled and tested. Never compiled in this image (no C# toolchain).
using System;
using System.Runtime.CompilerServices;
using UnityEngine;

namespace Unity.Mathematics
{
    public struct bool2 { public bool x, y; public bool2(bool x, bool y) { this.x = x; this.y = y; }
        public static bool2 operator |(bool2 a, bool2 b) => new bool2(a.x | b.x, a.y | b.y);
        public static bool2 operator &(bool2 a, bool2 b) => new bool2(a.x & b.x, a.y & b.y); }
    public struct bool3 { public bool x, y, z; public bool3(bool x, bool y, bool z) { this.x = x; this.y = y; this.z = z; }
        public static bool3 operator |(bool3 a, bool3 b) => new bool3(a.x | b.x, a.y | b.y, a.z | b.z);
        public static bool3 operator &(bool3 a, bool3 b) => new bool3(a.x & b.x, a.y & b.y, a.z & b.z); }

    public struct int2
    {
        public int x, y;
        public int2(int x, int y) { this.x = x; this.y = y; }
        public int2(int v) { x = v; y = v; }
        public int2(float2 v) { x = (int)v.x; y = (int)v.y; }
        public unsafe int this[int i] { get { fixed (int* p = &x) return p[i]; } set { fixed (int* p = &x) p[i] = value; } }
        public static implicit operator int2(int v) => new int2(v);
        public static int2 operator +(int2 a, int2 b) => new int2(a.x + b.x, a.y + b.y);
        public static int2 operator -(int2 a, int2 b) => new int2(a.x - b.x, a.y - b.y);
        public static int2 operator *(int2 a, int2 b) => new int2(a.x * b.x, a.y * b.y);
        public static int2 operator -(int2 a) => new int2(-a.x, -a.y);
        public static int2 operator ~(int2 a) => new int2(~a.x, ~a.y);
        public static int2 operator &(int2 a, int2 b) => new int2(a.x & b.x, a.y & b.y);
        public static int2 operator |(int2 a, int2 b) => new int2(a.x | b.x, a.y | b.y);
        public static int2 operator >>(int2 a, int s) => new int2(a.x >> s, a.y >> s);
        public static int2 operator <<(int2 a, int s) => new int2(a.x << s, a.y << s);
        public static bool2 operator <(int2 a, int2 b) => new bool2(a.x < b.x, a.y < b.y);
        public static bool2 operator >(int2 a, int2 b) => new bool2(a.x > b.x, a.y > b.y);
        public static bool2 operator <=(int2 a, int2 b) => new bool2(a.x <= b.x, a.y <= b.y);
        public static bool2 operator >=(int2 a, int2 b) => new bool2(a.x >= b.x, a.y >= b.y);
        public static bool2 operator ==(int2 a, int2 b) => new bool2(a.x == b.x, a.y == b.y);
        public static bool2 operator !=(int2 a, int2 b) => new bool2(a.x != b.x, a.y != b.y);
        public override bool Equals(object o) => o is int2 v && v.x == x && v.y == y;
        public override int GetHashCode() => x * 397 ^ y;
        public override string ToString() => $"int2({x}, {y})";
    }

    public struct int3
    {
        public int x, y, z;
        public int3(int x, int y, int z) { this.x = x; this.y = y; this.z = z; }
        public int3(int v) { x = v; y = v; z = v; }
        public int3(float3 v) { x = (int)v.x; y = (int)v.y; z = (int)v.z; }
        public int2 xz => new int2(x, z);
        public int2 xy => new int2(x, y);
        public static implicit operator int3(int v) => new int3(v);
        public static int3 operator +(int3 a, int3 b) => new int3(a.x + b.x, a.y + b.y, a.z + b.z);
        public static int3 operator -(int3 a, int3 b) => new int3(a.x - b.x, a.y - b.y, a.z - b.z);
        public static int3 operator *(int3 a, int3 b) => new int3(a.x * b.x, a.y * b.y, a.z * b.z);
        public static int3 operator >>(int3 a, int s) => new int3(a.x >> s, a.y >> s, a.z >> s);
        public override string ToString() => $"int3({x}, {y}, {z})";
    }

    public struct float2
    {
        public float x, y;
        public float2(float x, float y) { this.x = x; this.y = y; }
        public float2(float v) { x = v; y = v; }
        public unsafe float this[int i] { get { fixed (float* p = &x) return p[i]; } set { fixed (float* p = &x) p[i] = value; } }
        public float2 xy => this;
        public float2 yx => new float2(y, x);
        public float4 xyxy => new float4(x, y, x, y);
        public float4 xxyy => new float4(x, x, y, y);
        public static implicit operator float2(float v) => new float2(v);
        public static implicit operator float2(int2 v) => new float2(v.x, v.y);
        public static implicit operator float2(Vector2 v) => new float2(v.x, v.y);
        public static implicit operator Vector2(float2 v) => new Vector2(v.x, v.y);
        public static float2 operator +(float2 a, float2 b) => new float2(a.x + b.x, a.y + b.y);
        public static float2 operator -(float2 a, float2 b) => new float2(a.x - b.x, a.y - b.y);
        public static float2 operator *(float2 a, float2 b) => new float2(a.x * b.x, a.y * b.y);
        public static float2 operator /(float2 a, float2 b) => new float2(a.x / b.x, a.y / b.y);
        public static float2 operator +(float2 a, float b) => new float2(a.x + b, a.y + b);
        public static float2 operator -(float2 a, float b) => new float2(a.x - b, a.y - b);
        public static float2 operator *(float2 a, float b) => new float2(a.x * b, a.y * b);
        public static float2 operator /(float2 a, float b) => new float2(a.x / b, a.y / b);
        public static float2 operator +(float a, float2 b) => new float2(a + b.x, a + b.y);
        public static float2 operator -(float a, float2 b) => new float2(a - b.x, a - b.y);
        public static float2 operator *(float a, float2 b) => new float2(a * b.x, a * b.y);
        public static float2 operator /(float a, float2 b) => new float2(a / b.x, a / b.y);
        public static float2 operator -(float2 a) => new float2(-a.x, -a.y);
        public static bool2 operator <(float2 a, float2 b) => new bool2(a.x < b.x, a.y < b.y);
        public static bool2 operator >(float2 a, float2 b) => new bool2(a.x > b.x, a.y > b.y);
        public static bool2 operator <=(float2 a, float2 b) => new bool2(a.x <= b.x, a.y <= b.y);
        public static bool2 operator >=(float2 a, float2 b) => new bool2(a.x >= b.x, a.y >= b.y);
        public static bool2 operator <(float2 a, float b) => new bool2(a.x < b, a.y < b);
        public static bool2 operator >(float2 a, float b) => new bool2(a.x > b, a.y > b);
        public static bool2 operator <=(float2 a, float b) => new bool2(a.x <= b, a.y <= b);
        public static bool2 operator >=(float2 a, float b) => new bool2(a.x >= b, a.y >= b);
    }

    public struct float3
    {
        public float x, y, z;
        public float3(float x, float y, float z) { this.x = x; this.y = y; this.z = z; }
        public float3(float2 xy, float z) { x = xy.x; y = xy.y; this.z = z; }
        public float3(float v) { x = v; y = v; z = v; }
        public float2 xy => new float2(x, y);
        public float2 xz => new float2(x, z);
        public float3 yzx => new float3(y, z, x);
        public static implicit operator float3(float v) => new float3(v);
        public static implicit operator float3(int3 v) => new float3(v.x, v.y, v.z);
        public static implicit operator float3(Vector3 v) => new float3(v.x, v.y, v.z);
        public static implicit operator Vector3(float3 v) => new Vector3(v.x, v.y, v.z);
        public static float3 operator +(float3 a, float3 b) => new float3(a.x + b.x, a.y + b.y, a.z + b.z);
        public static float3 operator -(float3 a, float3 b) => new float3(a.x - b.x, a.y - b.y, a.z - b.z);
        public static float3 operator *(float3 a, float3 b) => new float3(a.x * b.x, a.y * b.y, a.z * b.z);
        public static float3 operator /(float3 a, float3 b) => new float3(a.x / b.x, a.y / b.y, a.z / b.z);
        public static float3 operator +(float3 a, float b) => new float3(a.x + b, a.y + b, a.z + b);
        public static float3 operator -(float3 a, float b) => new float3(a.x - b, a.y - b, a.z - b);
        public static float3 operator *(float3 a, float b) => new float3(a.x * b, a.y * b, a.z * b);
        public static float3 operator /(float3 a, float b) => new float3(a.x / b, a.y / b, a.z / b);
        public static float3 operator *(float a, float3 b) => new float3(a * b.x, a * b.y, a * b.z);
        public static float3 operator -(float3 a) => new float3(-a.x, -a.y, -a.z);
        public static bool3 operator <(float3 a, float3 b) => new bool3(a.x < b.x, a.y < b.y, a.z < b.z);
        public static bool3 operator >(float3 a, float3 b) => new bool3(a.x > b.x, a.y > b.y, a.z > b.z);
        public static bool3 operator <(float3 a, float b) => new bool3(a.x < b, a.y < b, a.z < b);
        public static bool3 operator >(float3 a, float b) => new bool3(a.x > b, a.y > b, a.z > b);
    }

    public struct float4
    {
        public float x, y, z, w;
        public float4(float x, float y, float z, float w) { this.x = x; this.y = y; this.z = z; this.w = w; }
        public float4(float3 v, float w) { x = v.x; y = v.y; z = v.z; this.w = w; }
        public float4(float2 a, float z, float w) { x = a.x; y = a.y; this.z = z; this.w = w; }
        public float4(float v) { x = v; y = v; z = v; w = v; }
        public unsafe float this[int i] { get { fixed (float* p = &x) return p[i]; } set { fixed (float* p = &x) p[i] = value; } }
        public float2 xy => new float2(x, y);
        public float2 zw => new float2(z, w);
        public float2 xz => new float2(x, z);
        public float3 xyz => new float3(x, y, z);
        public float3 xzw => new float3(x, z, w);
        public float3 yzw => new float3(y, z, w);
        public static implicit operator float4(float v) => new float4(v);
        public static float4 operator +(float4 a, float4 b) => new float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        public static float4 operator -(float4 a, float4 b) => new float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        public static float4 operator *(float4 a, float4 b) => new float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        public static float4 operator *(float4 a, float b) => new float4(a.x * b, a.y * b, a.z * b, a.w * b);
        public static float4 operator /(float4 a, float4 b) => new float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w);
    }

    public struct float4x4
    {
        public float4 c0, c1, c2, c3;
        public float4x4(float4 c0, float4 c1, float4 c2, float4 c3) { this.c0 = c0; this.c1 = c1; this.c2 = c2; this.c3 = c3; }
        public static float4x4 Scale(float x, float y, float z) => new float4x4(new float4(x, 0, 0, 0), new float4(0, y, 0, 0), new float4(0, 0, z, 0), new float4(0, 0, 0, 1));
        public static float4x4 Scale(float3 s) => Scale(s.x, s.y, s.z);
        public static float4x4 Translate(float3 t) => new float4x4(new float4(1, 0, 0, 0), new float4(0, 1, 0, 0), new float4(0, 0, 1, 0), new float4(t.x, t.y, t.z, 1));
        public static implicit operator float4x4(Matrix4x4 m) => new float4x4(
            new float4(m.m00, m.m10, m.m20, m.m30), new float4(m.m01, m.m11, m.m21, m.m31),
            new float4(m.m02, m.m12, m.m22, m.m32), new float4(m.m03, m.m13, m.m23, m.m33));
        public static implicit operator Matrix4x4(float4x4 f)
        {
            Matrix4x4 m = default;
            m.m00 = f.c0.x; m.m10 = f.c0.y; m.m20 = f.c0.z; m.m30 = f.c0.w;
            m.m01 = f.c1.x; m.m11 = f.c1.y; m.m21 = f.c1.z; m.m31 = f.c1.w;
            m.m02 = f.c2.x; m.m12 = f.c2.y; m.m22 = f.c2.z; m.m32 = f.c2.w;
            m.m03 = f.c3.x; m.m13 = f.c3.y; m.m23 = f.c3.z; m.m33 = f.c3.w;
            return m;
        }
    }

    public static class math
    {
        public static float2 float2(float x, float y) => new float2(x, y);
        public static float3 float3(float x, float y, float z) => new float3(x, y, z);
        public static float3 float3(float2 xy, float z) => new float3(xy, z);
        public static float4 float4(float x, float y, float z, float w) => new float4(x, y, z, w);
        public static float4 float4(float3 v, float w) => new float4(v, w);
        public static float4 float4(float2 a, float z, float w) => new float4(a, z, w);
        public static int2 int2(int x, int y) => new int2(x, y);
        public static int2 int2(float2 v) => new int2(v);
        public static int3 int3(int x, int y, int z) => new int3(x, y, z);
        public static int3 int3(float3 v) => new int3(v);

        public static float min(float x, float y) => float.IsNaN(y) || x < y ? x : y;
        public static float max(float x, float y) => float.IsNaN(y) || x > y ? x : y;
        public static int min(int x, int y) => x < y ? x : y;
        public static int max(int x, int y) => x > y ? x : y;
        public static float abs(float x) => BitConverter.Int32BitsToSingle(BitConverter.SingleToInt32Bits(x) & 0x7FFFFFFF);
        public static int abs(int x) => x < 0 ? -x : x;
        public static float floor(float x) => (float)Math.Floor((float)x);
        public static float ceil(float x) => (float)Math.Ceiling((float)x);
        public static float round(float x) => (float)Math.Round((float)x);   // half to even
        public static float frac(float x) => x - floor(x);
        public static float sign(float x) => (x > 0.0f ? 1.0f : 0.0f) - (x < 0.0f ? 1.0f : 0.0f);
        public static float sqrt(float x) => (float)Math.Sqrt((float)x);
        public static float rcp(float x) => 1.0f / x;
        public static float rsqrt(float x) => 1.0f / sqrt(x);
        public static float lerp(float a, float b, float t) => a + t * (b - a);
        public static float unlerp(float a, float b, float x) => (x - a) / (b - a);
        public static float select(float a, float b, bool c) => c ? b : a;
        public static int select(int a, int b, bool c) => c ? b : a;
        public static float clamp(float x, float a, float b) => max(a, min(b, x));
        public static int clamp(int x, int a, int b) => max(a, min(b, x));

        public static float2 min(float2 a, float2 b) => new float2(min(a.x, b.x), min(a.y, b.y));
        public static float2 max(float2 a, float2 b) => new float2(max(a.x, b.x), max(a.y, b.y));
        public static float3 min(float3 a, float3 b) => new float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
        public static float3 max(float3 a, float3 b) => new float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
        public static int3 min(int3 a, int3 b) => new int3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
        public static int3 max(int3 a, int3 b) => new int3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
        public static int3 clamp(int3 x, int3 a, int3 b) => max(a, min(b, x));
        public static float2 abs(float2 a) => new float2(abs(a.x), abs(a.y));
        public static float3 abs(float3 a) => new float3(abs(a.x), abs(a.y), abs(a.z));
        public static float2 floor(float2 a) => new float2(floor(a.x), floor(a.y));
        public static float3 floor(float3 a) => new float3(floor(a.x), floor(a.y), floor(a.z));
        public static float2 ceil(float2 a) => new float2(ceil(a.x), ceil(a.y));
        public static float3 ceil(float3 a) => new float3(ceil(a.x), ceil(a.y), ceil(a.z));
        public static float2 round(float2 a) => new float2(round(a.x), round(a.y));
        public static float2 frac(float2 a) => a - floor(a);
        public static float2 sign(float2 a) => new float2(sign(a.x), sign(a.y));
        public static float cmin(float2 a) => min(a.x, a.y);
        public static float cmax(float2 a) => max(a.x, a.y);
        public static float cmin(float3 a) => min(min(a.x, a.y), a.z);
        public static float cmax(float3 a) => max(max(a.x, a.y), a.z);
        public static int cmax(int3 a) => max(max(a.x, a.y), a.z);
        public static float dot(float2 a, float2 b) => a.x * b.x + a.y * b.y;
        public static float dot(float3 a, float3 b) => a.x * b.x + a.y * b.y + a.z * b.z;
        public static float3 cross(float3 x, float3 y) => (x * y.yzx - x.yzx * y).yzx;
        public static float2 normalize(float2 v) => rsqrt(dot(v, v)) * v;
        public static float3 normalize(float3 v) => rsqrt(dot(v, v)) * v;
        public static float2 lerp(float2 a, float2 b, float t) => a + t * (b - a);
        public static float3 lerp(float3 a, float3 b, float t) => a + t * (b - a);
        public static float2 select(float2 a, float2 b, bool c) => c ? b : a;
        public static bool any(bool2 b) => b.x || b.y;
        public static bool any(bool3 b) => b.x || b.y || b.z;
        public static bool all(bool2 b) => b.x && b.y;
        public static bool all(bool3 b) => b.x && b.y && b.z;
        public static float4 mul(float4x4 a, float4 b) => a.c0 * b.x + a.c1 * b.y + a.c2 * b.z + a.c3 * b.w;
        public static float4x4 mul(float4x4 a, float4x4 b) => new float4x4(mul(a, b.c0), mul(a, b.c1), mul(a, b.c2), mul(a, b.c3));

        // General inverse by cofactors in fp32 (the package orders its operations differently; host side only — Appendix A).
        public static float4x4 inverse(float4x4 mm)
        {
            float[] m = { mm.c0.x, mm.c0.y, mm.c0.z, mm.c0.w, mm.c1.x, mm.c1.y, mm.c1.z, mm.c1.w, mm.c2.x, mm.c2.y, mm.c2.z, mm.c2.w, mm.c3.x, mm.c3.y, mm.c3.z, mm.c3.w };
            float[] inv = new float[16];
            inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
            inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
            inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
            inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
            inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
            inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
            inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
            inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
            inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
            inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
            inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
            inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
            inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
            inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
            inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
            inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
            float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
            float rdet = 1.0f / det;
            for (int i = 0; i < 16; i++) inv[i] = inv[i] * rdet;
            return new float4x4(new float4(inv[0], inv[1], inv[2], inv[3]), new float4(inv[4], inv[5], inv[6], inv[7]),
                                new float4(inv[8], inv[9], inv[10], inv[11]), new float4(inv[12], inv[13], inv[14], inv[15]));
        }
    }
}
