// unity_shim.hpp — TEST INFRASTRUCTURE ONLY (part of the oracle/_ref recipe).
//
// The reference (pipliz/cpuvox) is C# on Unity.Mathematics 1.2.6 / Unity.Collections 1.2.4 / Unity.Jobs / UnityEngine
// 2022.3; none of those packages is vendored under /root/reference and the image has no C# toolchain. oracle/refbuild/cs2cpp.py
// rewrites the reference's own .cs files (read where they lie, never copied into the repository) into C++ token for token;
// this header supplies, in C++, the subset of the Unity API those files call. Everything here is OUR restatement of the
// published behaviour of those packages (SURVEY.md Appendix A lists the assumptions); the algorithm that runs on top of it
// is the reference's own source text.
//
// Arithmetic: IEEE fp32, built with -ffp-contract=off -fno-fast-math, i.e. what RyuJIT emits for C# float code on x64
// (SSE scalar, no FMA contraction). The shipping player is Burst FloatMode.Fast, which neither this nor any strict build matches.
#pragma once
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <alloca.h>

namespace cpuvox_ref {

typedef uint8_t byte;
typedef int8_t sbyte;
typedef uint16_t ushort;
typedef uint32_t uint;
typedef std::string string;

// ---------------------------------------------------------------------------------------------------------------------
// C# language helpers
// ---------------------------------------------------------------------------------------------------------------------
static const float cs_float_Epsilon = 1.401298464324817e-45f;  // float.Epsilon: the smallest positive denormal
static const float cs_float_NegativeInfinity = -INFINITY;
static const float cs_float_PositiveInfinity = INFINITY;
static const int cs_int_MaxValue = INT_MAX;
static const int cs_int_MinValue = INT_MIN;

// stackalloc T[n] (Burst zero-initialises it, DrawSegmentRayJob.cs:208 relies on that)
#define cs_stackalloc(T, n) ((T*)memset(alloca(sizeof(T) * (size_t)(n)), 0, sizeof(T) * (size_t)(n)))

// (int)someFloat on x64 (.NET and Burst): cvttss2si, out-of-range and NaN give int.MinValue
inline int cs_f2i(float f) {
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
}
inline int cs_f2i(double f) {
    if (!(f >= -2147483648.0 && f < 2147483648.0)) return INT_MIN;
    return (int)f;
}
inline int cs_f2i(int i) { return i; }
inline int cs_f2i(long i) { return (int)i; }
inline int cs_f2i(short i) { return i; }
inline int cs_f2i(ushort i) { return i; }
inline int cs_f2i(byte i) { return i; }

// object initialiser: new T { a = x, b = y } / new T(args) { a = x }
template <class T, class F>
inline T cs_init(T obj, F&& f) {
    f(obj);
    return obj;
}

// `new SomeClass(args)`: objects of the few reference types the translation meets are held by value (their pixel / vertex
// stores are shared handles), so a fresh object is just the temporary
template <class T>
inline T cs_new(T&& obj) { return std::move(obj); }

struct cs_exception : std::runtime_error {
    explicit cs_exception(const char* what) : std::runtime_error(what) {}
};
#define CS_EXCEPTION(Name)                                      \
    struct Name : cs_exception {                                \
        Name() : cs_exception(#Name) {}                         \
        explicit Name(const char* m) : cs_exception(m) {}       \
    }
CS_EXCEPTION(InvalidOperationException);
CS_EXCEPTION(ArgumentOutOfRangeException);
CS_EXCEPTION(ArgumentException);
CS_EXCEPTION(OutOfMemoryException);

// T[] : a managed array is a reference; copies share the storage
template <class T>
struct ManagedArray {
    std::shared_ptr<std::vector<T>> store;
    int Length = 0;
    ManagedArray() {}
    ManagedArray(std::nullptr_t) {}
    explicit ManagedArray(int n) : store(std::make_shared<std::vector<T>>((size_t)n)), Length(n) {}
    ManagedArray(std::initializer_list<T> l) : store(std::make_shared<std::vector<T>>(l)), Length((int)l.size()) {}
    T& operator[](long i) const {
        if (i < 0 || i >= Length) throw cs_exception("IndexOutOfRangeException");
        return (*store)[(size_t)i];
    }
    T* data() const { return store ? store->data() : nullptr; }
    bool operator==(std::nullptr_t) const { return !store; }
    bool operator!=(std::nullptr_t) const { return (bool)store; }
    T* begin() const { return data(); }
    T* end() const { return data() + Length; }
};

// List<T>
template <class T>
struct List {
    std::shared_ptr<std::vector<T>> store;
    List() {}
    List(std::nullptr_t) {}
    static List Create() {
        List l;
        l.store = std::make_shared<std::vector<T>>();
        return l;
    }
    struct CountProxy {
        const List* l;
        operator int() const { return (int)l->store->size(); }
    };
    int Count() const { return (int)store->size(); }
    void Add(const T& t) { store->push_back(t); }
    void Clear() { store->clear(); }
    T& operator[](long i) const { return (*store)[(size_t)i]; }
    bool operator==(std::nullptr_t) const { return !store; }
    bool operator!=(std::nullptr_t) const { return (bool)store; }
    template <class F>
    void Sort(F cmp) {
        // List<T>.Sort is an unstable introsort; every use in the reference is followed by an order-independent merge of equal keys.
        std::sort(store->begin(), store->end(), [&](const T& a, const T& b) { return cmp(a, b) < 0; });
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Unity.Mathematics 1.2.6 (subset). Vectors are plain structs; float4x4 is column major (c0..c3).
// ---------------------------------------------------------------------------------------------------------------------
struct bool2 { bool x, y; bool2() : x(false), y(false) {} bool2(bool x, bool y) : x(x), y(y) {} bool2(bool v) : x(v), y(v) {} };
struct bool3 { bool x, y, z; bool3() : x(false), y(false), z(false) {} bool3(bool x, bool y, bool z) : x(x), y(y), z(z) {} bool3(bool v) : x(v), y(v), z(v) {} };
struct bool4 { bool x, y, z, w; bool4() : x(false), y(false), z(false), w(false) {} bool4(bool x, bool y, bool z, bool w) : x(x), y(y), z(z), w(w) {} };
inline bool2 operator|(bool2 a, bool2 b) { return bool2(a.x | b.x, a.y | b.y); }
inline bool2 operator&(bool2 a, bool2 b) { return bool2(a.x & b.x, a.y & b.y); }
inline bool3 operator|(bool3 a, bool3 b) { return bool3(a.x | b.x, a.y | b.y, a.z | b.z); }
inline bool3 operator&(bool3 a, bool3 b) { return bool3(a.x & b.x, a.y & b.y, a.z & b.z); }
inline bool any(bool2 b) { return b.x || b.y; }
inline bool any(bool3 b) { return b.x || b.y || b.z; }
inline bool all(bool2 b) { return b.x && b.y; }
inline bool all(bool3 b) { return b.x && b.y && b.z; }

struct float2;
struct float3;
struct float4;
struct Vector2;
struct Vector3;

struct int2 {
    int x, y;
    int2() : x(0), y(0) {}
    int2(int x, int y) : x(x), y(y) {}
    int2(int v) : x(v), y(v) {}
    explicit int2(const float2& f);
    int& operator[](int i) { return (&x)[i]; }
    int operator[](int i) const { return (&x)[i]; }
    int2 xy() const { return *this; }
    int2 yx() const { return int2(y, x); }
    int2& operator+=(int2 b) { x += b.x; y += b.y; return *this; }
    int2& operator-=(int2 b) { x -= b.x; y -= b.y; return *this; }
    int2& operator*=(int2 b) { x *= b.x; y *= b.y; return *this; }
    int2& operator&=(int2 b) { x &= b.x; y &= b.y; return *this; }
    int2& operator>>=(int s) { x >>= s; y >>= s; return *this; }
    int2& operator<<=(int s) { x <<= s; y <<= s; return *this; }
};
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); }
inline int2 operator*(int2 a, int2 b) { return int2(a.x * b.x, a.y * b.y); }
inline int2 operator&(int2 a, int2 b) { return int2(a.x & b.x, a.y & b.y); }
inline int2 operator|(int2 a, int2 b) { return int2(a.x | b.x, a.y | b.y); }
inline int2 operator~(int2 a) { return int2(~a.x, ~a.y); }
inline int2 operator-(int2 a) { return int2(-a.x, -a.y); }
inline int2 operator>>(int2 a, int s) { return int2(a.x >> s, a.y >> s); }
inline int2 operator<<(int2 a, int s) { return int2(a.x << s, a.y << s); }
inline bool2 operator<(int2 a, int2 b) { return bool2(a.x < b.x, a.y < b.y); }
inline bool2 operator>(int2 a, int2 b) { return bool2(a.x > b.x, a.y > b.y); }
inline bool2 operator<=(int2 a, int2 b) { return bool2(a.x <= b.x, a.y <= b.y); }
inline bool2 operator>=(int2 a, int2 b) { return bool2(a.x >= b.x, a.y >= b.y); }
inline bool2 operator==(int2 a, int2 b) { return bool2(a.x == b.x, a.y == b.y); }
inline bool2 operator!=(int2 a, int2 b) { return bool2(a.x != b.x, a.y != b.y); }

struct int3 {
    int x, y, z;
    int3() : x(0), y(0), z(0) {}
    int3(int x, int y, int z) : x(x), y(y), z(z) {}
    int3(int v) : x(v), y(v), z(v) {}
    explicit int3(const float3& f);
    int& operator[](int i) { return (&x)[i]; }
    int operator[](int i) const { return (&x)[i]; }
    int2 xz() const { return int2(x, z); }
    int2 xy() const { return int2(x, y); }
};
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline int3 operator-(int3 a, int3 b) { return int3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline int3 operator*(int3 a, int3 b) { return int3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline int3 operator>>(int3 a, int s) { return int3(a.x >> s, a.y >> s, a.z >> s); }
inline bool3 operator<(int3 a, int3 b) { return bool3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bool3 operator>(int3 a, int3 b) { return bool3(a.x > b.x, a.y > b.y, a.z > b.z); }

struct float4;
struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float x, float y) : x(x), y(y) {}
    float2(float v) : x(v), y(v) {}
    float2(int2 v) : x((float)v.x), y((float)v.y) {}   // implicit int2 -> float2
    float2(const Vector2& v);
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    float2 xy() const { return *this; }
    float2 yx() const { return float2(y, x); }
    float4 xyxy() const;
    float4 xxyy() const;
    float2& operator+=(float2 b) { x += b.x; y += b.y; return *this; }
    float2& operator-=(float2 b) { x -= b.x; y -= b.y; return *this; }
    float2& operator*=(float2 b) { x *= b.x; y *= b.y; return *this; }
    float2& operator/=(float2 b) { x /= b.x; y /= b.y; return *this; }
};
inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(float2 a, float2 b) { return float2(a.x - b.x, a.y - b.y); }
inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
inline float2 operator/(float2 a, float2 b) { return float2(a.x / b.x, a.y / b.y); }
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline bool2 operator<(float2 a, float2 b) { return bool2(a.x < b.x, a.y < b.y); }
inline bool2 operator>(float2 a, float2 b) { return bool2(a.x > b.x, a.y > b.y); }
inline bool2 operator<=(float2 a, float2 b) { return bool2(a.x <= b.x, a.y <= b.y); }
inline bool2 operator>=(float2 a, float2 b) { return bool2(a.x >= b.x, a.y >= b.y); }
inline int2::int2(const float2& f) : x(cs_f2i(f.x)), y(cs_f2i(f.y)) {}

struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float x, float y, float z) : x(x), y(y), z(z) {}
    float3(float2 xy, float z) : x(xy.x), y(xy.y), z(z) {}
    float3(float v) : x(v), y(v), z(v) {}
    float3(int3 v) : x((float)v.x), y((float)v.y), z((float)v.z) {}  // implicit int3 -> float3
    float3(const Vector3& v);
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    float2 xy() const { return float2(x, y); }
    float2 xz() const { return float2(x, z); }
    float2 yz() const { return float2(y, z); }
    float3 yzx() const { return float3(y, z, x); }
    float3& operator+=(float3 b) { x += b.x; y += b.y; z += b.z; return *this; }
    float3& operator-=(float3 b) { x -= b.x; y -= b.y; z -= b.z; return *this; }
    float3& operator*=(float3 b) { x *= b.x; y *= b.y; z *= b.z; return *this; }
};
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float3 b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline bool3 operator<(float3 a, float3 b) { return bool3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bool3 operator>(float3 a, float3 b) { return bool3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline int3::int3(const float3& f) : x(cs_f2i(f.x)), y(cs_f2i(f.y)), z(cs_f2i(f.z)) {}

struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float x, float y, float z, float w) : x(x), y(y), z(z), w(w) {}
    float4(float3 v, float w) : x(v.x), y(v.y), z(v.z), w(w) {}
    float4(float2 a, float z, float w) : x(a.x), y(a.y), z(z), w(w) {}
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    float4(float v) : x(v), y(v), z(v), w(v) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    float2 xy() const { return float2(x, y); }
    float2 zw() const { return float2(z, w); }
    float2 xz() const { return float2(x, z); }
    float3 xyz() const { return float3(x, y, z); }
    float3 xzw() const { return float3(x, z, w); }
    float3 yzw() const { return float3(y, z, w); }
};
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float4 operator*(float4 a, float4 b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline float4 operator/(float4 a, float4 b) { return float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline float4 float2::xyxy() const { return float4(x, y, x, y); }
inline float4 float2::xxyy() const { return float4(x, x, y, y); }

inline float2 operator+(float2 a, float b) { return a + float2(b); }
inline float2 operator+(float a, float2 b) { return float2(a) + b; }
inline float2 operator-(float2 a, float b) { return a - float2(b); }
inline float2 operator-(float a, float2 b) { return float2(a) - b; }
inline float2 operator*(float2 a, float b) { return a * float2(b); }
inline float2 operator*(float a, float2 b) { return float2(a) * b; }
inline float2 operator/(float2 a, float b) { return a / float2(b); }
inline float2 operator/(float a, float2 b) { return float2(a) / b; }
inline float3 operator+(float3 a, float b) { return a + float3(b); }
inline float3 operator+(float a, float3 b) { return float3(a) + b; }
inline float3 operator-(float3 a, float b) { return a - float3(b); }
inline float3 operator-(float a, float3 b) { return float3(a) - b; }
inline float3 operator*(float3 a, float b) { return a * float3(b); }
inline float3 operator*(float a, float3 b) { return float3(a) * b; }
inline float3 operator/(float3 a, float b) { return a / float3(b); }
inline float3 operator/(float a, float3 b) { return float3(a) / b; }
inline float4 operator+(float4 a, float b) { return a + float4(b); }
inline float4 operator+(float a, float4 b) { return float4(a) + b; }
inline float4 operator-(float4 a, float b) { return a - float4(b); }
inline float4 operator-(float a, float4 b) { return float4(a) - b; }
inline float4 operator*(float4 a, float b) { return a * float4(b); }
inline float4 operator*(float a, float4 b) { return float4(a) * b; }
inline float4 operator/(float4 a, float b) { return a / float4(b); }
inline float4 operator/(float a, float4 b) { return float4(a) / b; }
inline int2 operator+(int2 a, int b) { return a + int2(b); }
inline int2 operator+(int a, int2 b) { return int2(a) + b; }
inline int2 operator-(int2 a, int b) { return a - int2(b); }
inline int2 operator-(int a, int2 b) { return int2(a) - b; }
inline int2 operator*(int2 a, int b) { return a * int2(b); }
inline int2 operator*(int a, int2 b) { return int2(a) * b; }
inline int3 operator+(int3 a, int b) { return a + int3(b); }
inline int3 operator+(int a, int3 b) { return int3(a) + b; }
inline int3 operator-(int3 a, int b) { return a - int3(b); }
inline int3 operator-(int a, int3 b) { return int3(a) - b; }
inline int3 operator*(int3 a, int b) { return a * int3(b); }
inline int3 operator*(int a, int3 b) { return int3(a) * b; }
inline bool2 operator<(float2 a, float b) { return a < float2(b); }
inline bool2 operator>(float2 a, float b) { return a > float2(b); }
inline bool2 operator<=(float2 a, float b) { return a <= float2(b); }
inline bool2 operator>=(float2 a, float b) { return a >= float2(b); }
inline bool3 operator<(float3 a, float b) { return a < float3(b); }
inline bool3 operator>(float3 a, float b) { return a > float3(b); }
inline bool2 operator<(int2 a, int b) { return a < int2(b); }
inline bool2 operator>(int2 a, int b) { return a > int2(b); }
inline bool2 operator<=(int2 a, int b) { return a <= int2(b); }
inline bool2 operator>=(int2 a, int b) { return a >= int2(b); }
inline int2 operator&(int2 a, int b) { return a & int2(b); }
inline float2 operator*(float2 a, int b) { return a * float2((float)b); }
struct Matrix4x4;
struct float4x4 {
    float4 c0, c1, c2, c3;
    float4x4() {}
    float4x4(float4 c0, float4 c1, float4 c2, float4 c3) : c0(c0), c1(c1), c2(c2), c3(c3) {}
    float4x4(const Matrix4x4& m);  // implicit Matrix4x4 -> float4x4 (column for column)
    static float4x4 identity() { return float4x4(float4(1, 0, 0, 0), float4(0, 1, 0, 0), float4(0, 0, 1, 0), float4(0, 0, 0, 1)); }
    static float4x4 Scale(float x, float y, float z) { return float4x4(float4(x, 0, 0, 0), float4(0, y, 0, 0), float4(0, 0, z, 0), float4(0, 0, 0, 1)); }
    static float4x4 Scale(float3 s) { return Scale(s.x, s.y, s.z); }
    static float4x4 Translate(float3 t) { return float4x4(float4(1, 0, 0, 0), float4(0, 1, 0, 0), float4(0, 0, 1, 0), float4(t.x, t.y, t.z, 1)); }
};

// scalar math.* (the reference uses `using static Unity.Mathematics.math`)
inline float min(float x, float y) { return (std::isnan(y) || x < y) ? x : y; }
inline float max(float x, float y) { return (std::isnan(y) || x > y) ? x : y; }
inline int min(int x, int y) { return x < y ? x : y; }
inline int max(int x, int y) { return x > y ? x : y; }
inline float abs(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0x7FFFFFFFu; memcpy(&x, &u, 4); return x; }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return (float)std::floor((double)x); }
inline float ceil(float x) { return (float)std::ceil((double)x); }
inline double floor(double x) { return std::floor(x); }
inline double ceil(double x) { return std::ceil(x); }
inline float round(float x) { return (float)std::nearbyint((double)x); }  // System.Math.Round: half to even
inline float frac(float x) { return x - floor(x); }
inline float sign(float x) { return (x > 0.0f ? 1.0f : 0.0f) - (x < 0.0f ? 1.0f : 0.0f); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float rcp(float x) { return 1.0f / x; }
inline float rsqrt(float x) { return 1.0f / sqrt(x); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float unlerp(float a, float b, float x) { return (x - a) / (b - a); }
inline float select(float a, float b, bool c) { return c ? b : a; }
inline int select(int a, int b, bool c) { return c ? b : a; }
inline float clamp(float x, float a, float b) { return max(a, min(b, x)); }
inline int clamp(int x, int a, int b) { return max(a, min(b, x)); }

// vector math.*
inline float2 min(float2 a, float2 b) { return float2(min(a.x, b.x), min(a.y, b.y)); }
inline float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline int3 min(int3 a, int3 b) { return int3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline int3 max(int3 a, int3 b) { return int3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline int3 clamp(int3 x, int3 a, int3 b) { return max(a, min(b, x)); }
inline float2 abs(float2 a) { return float2(abs(a.x), abs(a.y)); }
inline float3 abs(float3 a) { return float3(abs(a.x), abs(a.y), abs(a.z)); }
inline float2 floor(float2 a) { return float2(floor(a.x), floor(a.y)); }
inline float3 floor(float3 a) { return float3(floor(a.x), floor(a.y), floor(a.z)); }
inline float2 ceil(float2 a) { return float2(ceil(a.x), ceil(a.y)); }
inline float3 ceil(float3 a) { return float3(ceil(a.x), ceil(a.y), ceil(a.z)); }
inline float2 round(float2 a) { return float2(round(a.x), round(a.y)); }
inline float2 frac(float2 a) { return a - floor(a); }
inline float2 sign(float2 a) { return float2(sign(a.x), sign(a.y)); }
inline float cmin(float2 a) { return min(a.x, a.y); }
inline float cmax(float2 a) { return max(a.x, a.y); }
inline float cmin(float3 a) { return min(min(a.x, a.y), a.z); }
inline float cmax(float3 a) { return max(max(a.x, a.y), a.z); }
inline int cmax(int3 a) { return max(max(a.x, a.y), a.z); }
inline int cmin(int3 a) { return min(min(a.x, a.y), a.z); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 x, float3 y) { return (x * y.yzx() - x.yzx() * y).yzx(); }
inline float2 normalize(float2 v) { return float2(rsqrt(dot(v, v))) * v; }
inline float3 normalize(float3 v) { return float3(rsqrt(dot(v, v))) * v; }
inline float2 lerp(float2 a, float2 b, float t) { return a + float2(t) * (b - a); }
inline float3 lerp(float3 a, float3 b, float t) { return a + float3(t) * (b - a); }
inline float2 select(float2 a, float2 b, bool c) { return c ? b : a; }
inline float4 mul(const float4x4& a, float4 b) { return a.c0 * float4(b.x) + a.c1 * float4(b.y) + a.c2 * float4(b.z) + a.c3 * float4(b.w); }
inline float4x4 mul(const float4x4& a, const float4x4& b) { return float4x4(mul(a, b.c0), mul(a, b.c1), mul(a, b.c2), mul(a, b.c3)); }
// General inverse by cofactors, fp32. Unity.Mathematics' inverse(float4x4) orders its operations differently (SIMD shuffles);
// host side only (segment setup), stated as an assumption in SURVEY.md Appendix A.
inline float4x4 inverse(const float4x4& mm) {
    float m[16], inv[16];
    const float4* cols[4] = {&mm.c0, &mm.c1, &mm.c2, &mm.c3};
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) m[j * 4 + i] = (*cols[j])[i];
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
    float4x4 r;
    float4* rc[4] = {&r.c0, &r.c1, &r.c2, &r.c3};
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) (*rc[j])[i] = inv[j * 4 + i] * rdet;
    return r;
}
struct math {  // explicit `math.ceil(...)` call sites
    static float ceil(float x) { return cpuvox_ref::ceil(x); }
    static float floor(float x) { return cpuvox_ref::floor(x); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Unity.Collections / Unity.Jobs / UnsafeUtility (subset)
// ---------------------------------------------------------------------------------------------------------------------
enum class Allocator { Invalid, None, Temp, TempJob, Persistent };
enum class NativeArrayOptions { UninitializedMemory, ClearMemory };

struct UnsafeUtility {
    static void* Malloc(long bytes, int align, Allocator) {
        void* p = nullptr;
        if (posix_memalign(&p, (size_t)std::max(align, (int)sizeof(void*)), (size_t)std::max<long>(bytes, 1)) != 0) throw OutOfMemoryException();
        return p;
    }
    static void Free(void* p, Allocator) { free(p); }
    static void MemClear(void* p, long bytes) { memset(p, 0, (size_t)bytes); }
    static void MemCpy(void* dst, const void* src, long bytes) { memcpy(dst, src, (size_t)bytes); }
    template <class T> static int SizeOf() { return (int)sizeof(T); }
    template <class T> static int AlignOf() { return (int)alignof(T); }
    template <class T> static void CopyStructureToPtr(T& s, void* p) { memcpy(p, &s, sizeof(T)); }
    template <class T> static void CopyPtrToStructure(void* p, T& s) { memcpy(&s, p, sizeof(T)); }
};

// NativeArray<T>: a handle to unmanaged memory (copies alias)
template <class T>
struct NativeArray {
    T* ptr = nullptr;
    int Length = 0;
    NativeArray() {}
    NativeArray(int n, Allocator a, NativeArrayOptions o = NativeArrayOptions::ClearMemory) : Length(n) {
        ptr = (T*)UnsafeUtility::Malloc((long)sizeof(T) * n, (int)alignof(T), a);
        if (o == NativeArrayOptions::ClearMemory) memset((void*)ptr, 0, sizeof(T) * (size_t)n);
    }
    T& operator[](long i) const {
        if (i < 0 || i >= Length) throw cs_exception("IndexOutOfRangeException (NativeArray)");
        return ptr[i];
    }
    void* GetUnsafePtr() const { return ptr; }
    void* GetUnsafeReadOnlyPtr() const { return ptr; }
    void Dispose() { free((void*)ptr); ptr = nullptr; Length = 0; }
    static void Copy(const NativeArray& src, NativeArray& dst, int n) { memcpy((void*)dst.ptr, (const void*)src.ptr, sizeof(T) * (size_t)n); }
};

// NativeList<T>: handle; Length is live (RenderJob reads it after TraceToFirstColumnJob appended through the ParallelWriter)
template <class T>
struct NativeList {
    struct Header {
        T* data;
        std::atomic<int> length;
        int capacity;
    };
    Header* h = nullptr;
    struct LengthProxy {
        Header* const* hh;
        operator int() const { return (*hh)->length.load(); }
    };
    LengthProxy Length{&h};
    NativeList() {}
    NativeList(int capacity, Allocator a) {
        h = new Header();
        h->data = (T*)UnsafeUtility::Malloc((long)sizeof(T) * std::max(capacity, 1), (int)alignof(T), a);
        h->length = 0;
        h->capacity = capacity;
    }
    NativeList(const NativeList& o) : h(o.h), Length{&h} {}
    NativeList& operator=(const NativeList& o) { h = o.h; return *this; }
    T& operator[](long i) const {
        if (i < 0 || i >= h->length.load()) throw cs_exception("IndexOutOfRangeException (NativeList)");
        return h->data[i];
    }
    struct ParallelWriter {
        Header* h = nullptr;
        void AddNoResize(const T& v) {
            int i = h->length.fetch_add(1);
            if (i >= h->capacity) throw cs_exception("NativeList.AddNoResize over capacity");
            h->data[i] = v;
        }
    };
    ParallelWriter AsParallelWriter() const { ParallelWriter w; w.h = h; return w; }
    void Dispose() { if (h) { free((void*)h->data); delete h; h = nullptr; } }
};

// Jobs: Schedule runs the job to completion before returning (dependencies are expressed by call order in the reference),
// batches are handed to worker threads through a shared counter, like Unity's IJobParallelFor work stealing.
struct JobHandle { void Complete() {} };
extern int g_job_threads;  // 0 / 1: inline
template <class Job>
struct IJobParallelFor {
    JobHandle Schedule(int arrayLength, int innerloopBatchCount, JobHandle = JobHandle()) {
        Job& job = *static_cast<Job*>(this);
        int threads = g_job_threads;
        if (threads <= 1 || arrayLength <= innerloopBatchCount) {
            for (int i = 0; i < arrayLength; i++) job.Execute(i);
            return JobHandle();
        }
        std::atomic<int> next(0);
        std::atomic<bool> failed(false);
        std::string err;
        auto worker = [&]() {
            try {
                for (;;) {
                    int b = next.fetch_add(innerloopBatchCount);
                    if (b >= arrayLength) break;
                    int e = std::min(arrayLength, b + innerloopBatchCount);
                    for (int i = b; i < e; i++) job.Execute(i);
                }
            } catch (const std::exception& ex) {
                if (!failed.exchange(true)) err = ex.what();
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; t++) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
        if (failed) throw cs_exception(strdup(err.c_str()));
        return JobHandle();
    }
};
struct IDisposable {};

struct SpinLock {
    SpinLock(bool = false) {}
    void Enter(bool& taken) { taken = true; }
    void Exit() {}
};
struct Interlocked {
    static int Add(int& location, int v) { location += v; return location; }
    template <class T, class U> static T CompareExchange(T& location, const T& value, U comparand) {
        T old = location;
        if (old == comparand) location = value;
        return old;
    }
};
struct Environment { static const int ProcessorCount = 1; };
struct ParallelOptions { int MaxDegreeOfParallelism = 1; };
struct Parallel {
    template <class F> static void For(int from, int to, F f) { for (int i = from; i < to; i++) f(i); }
    template <class F> static void For(int from, int to, const ParallelOptions&, F f) { for (int i = from; i < to; i++) f(i); }
};

// ---------------------------------------------------------------------------------------------------------------------
// UnityEngine (subset; closed source — behaviour as documented)
// ---------------------------------------------------------------------------------------------------------------------
struct Mathf {
    static constexpr float Deg2Rad = 0.0174532924f;  // (float)(PI * 2 / 360)
    static constexpr float Rad2Deg = 57.29578f;
    static int RoundToInt(float f) { return cs_f2i(std::nearbyint((double)f)); }  // (int)Math.Round(f): half to even
    static float Sin(float f) { return (float)std::sin((double)f); }
    static float Cos(float f) { return (float)std::cos((double)f); }
    static float Tan(float f) { return (float)std::tan((double)f); }
    static float Acos(float f) { return (float)std::acos((double)f); }
    static float Sqrt(float f) { return (float)std::sqrt((double)f); }
    static float Abs(float f) { return std::fabs(f); }
    static int Abs(int f) { return f < 0 ? -f : f; }
    static float Sign(float f) { return f >= 0.0f ? 1.0f : -1.0f; }
    static int Max(int a, int b) { return a > b ? a : b; }
    static int Min(int a, int b) { return a < b ? a : b; }
    static float Max(float a, float b) { return a > b ? a : b; }
    static float Min(float a, float b) { return a < b ? a : b; }
    static float Clamp(float v, float a, float b) { return v < a ? a : (v > b ? b : v); }
    static int NextPowerOfTwo(int v) {
        v -= 1; v |= v >> 16; v |= v >> 8; v |= v >> 4; v |= v >> 2; v |= v >> 1;
        return v + 1;
    }
};

struct Vector2 {
    float x, y;
    Vector2() : x(0), y(0) {}
    Vector2(float x, float y) : x(x), y(y) {}
    Vector2(const float2& f) : x(f.x), y(f.y) {}
    // Vector2.SignedAngle(from, to) = Angle(from, to) * Sign(from.x * to.y - from.y * to.x)
    static float Angle(Vector2 from, Vector2 to) {
        float denominator = (float)std::sqrt((double)((from.x * from.x + from.y * from.y) * (to.x * to.x + to.y * to.y)));
        if (denominator < 1e-15f) return 0.0f;
        float d = Mathf::Clamp((from.x * to.x + from.y * to.y) / denominator, -1.0f, 1.0f);
        return (float)std::acos((double)d) * Mathf::Rad2Deg;
    }
    static float SignedAngle(Vector2 from, Vector2 to) {
        float unsigned_angle = Angle(from, to);
        float sign = Mathf::Sign(from.x * to.y - from.y * to.x);
        return unsigned_angle * sign;
    }
};
inline float2::float2(const Vector2& v) : x(v.x), y(v.y) {}

struct Vector3 {
    float x, y, z;
    Vector3() : x(0), y(0), z(0) {}
    Vector3(float x, float y, float z) : x(x), y(y), z(z) {}
    Vector3(const float3& f) : x(f.x), y(f.y), z(f.z) {}
    static Vector3 zero;
    static Vector3 one;
    static Vector3 up;
    static Vector3 forward;
    static float Dot(Vector3 a, Vector3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    static Vector3 Cross(Vector3 a, Vector3 b) { return Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
    static float Magnitude(Vector3 a) { return (float)std::sqrt((double)(a.x * a.x + a.y * a.y + a.z * a.z)); }
    static Vector3 Normalize(Vector3 a) {
        float mag = Magnitude(a);
        if (mag > 1e-05f) return Vector3(a.x / mag, a.y / mag, a.z / mag);
        return Vector3(0, 0, 0);
    }
    static float Distance(Vector3 a, Vector3 b) {
        float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        return (float)std::sqrt((double)(dx * dx + dy * dy + dz * dz));
    }
};
inline Vector3 operator+(Vector3 a, Vector3 b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(Vector3 a, Vector3 b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator*(Vector3 a, float d) { return Vector3(a.x * d, a.y * d, a.z * d); }
inline Vector3 operator*(float d, Vector3 a) { return Vector3(a.x * d, a.y * d, a.z * d); }
inline float3::float3(const Vector3& v) : x(v.x), y(v.y), z(v.z) {}

struct Vector4 {
    float x, y, z, w;
    Vector4() : x(0), y(0), z(0), w(0) {}
    Vector4(float x, float y, float z, float w) : x(x), y(y), z(z), w(w) {}
    float operator[](int i) const { return (&x)[i]; }
};

struct Color32 {
    byte r, g, b, a;
    Color32() : r(0), g(0), b(0), a(0) {}
    Color32(byte r, byte g, byte b, byte a) : r(r), g(g), b(b), a(a) {}
};
struct Color {
    float r, g, b, a;
    Color() : r(0), g(0), b(0), a(0) {}
    Color(float r, float g, float b, float a = 1.0f) : r(r), g(g), b(b), a(a) {}
    Color(const Color32& c) : r(c.r / 255.0f), g(c.g / 255.0f), b(c.b / 255.0f), a(c.a / 255.0f) {}  // implicit Color32 -> Color
    operator Color32() const {  // implicit Color -> Color32: (byte)Math.Round(Clamp01(c) * 255f)
        auto cv = [](float v) { v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); return (byte)std::nearbyint((double)(v * 255.0f)); };
        return Color32(cv(r), cv(g), cv(b), cv(a));
    }
    static Color red;
};
inline Color operator*(Color a, Color b) { return Color(a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a); }

struct Quaternion {
    float x, y, z, w;
    Quaternion() : x(0), y(0), z(0), w(1) {}
    Quaternion(float x, float y, float z, float w) : x(x), y(y), z(z), w(w) {}
};
inline Quaternion operator*(Quaternion a, Quaternion b) {
    return Quaternion(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                      a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
inline Vector3 operator*(Quaternion q, Vector3 p) {
    float x2 = q.x * 2.0f, y2 = q.y * 2.0f, z2 = q.z * 2.0f;
    float xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
    float xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
    float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    Vector3 r;
    r.x = (1.0f - (yy + zz)) * p.x + (xy - wz) * p.y + (xz + wy) * p.z;
    r.y = (xy + wz) * p.x + (1.0f - (xx + zz)) * p.y + (yz - wx) * p.z;
    r.z = (xz - wy) * p.x + (yz + wx) * p.y + (1.0f - (xx + yy)) * p.z;
    return r;
}

struct Matrix4x4 {
    float m[4][4];  // m[col][row], column major like Unity's m00..m33 storage
    Matrix4x4() { memset(m, 0, sizeof m); }
    Matrix4x4(const float4x4& f) {
        const float4* c[4] = {&f.c0, &f.c1, &f.c2, &f.c3};
        for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) m[j][i] = (*c[j])[i];
    }
    static Matrix4x4 identity_() { Matrix4x4 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f; return r; }
    static Matrix4x4 identity;
    static Matrix4x4 Scale(Vector3 s) { Matrix4x4 r = identity_(); r.m[0][0] = s.x; r.m[1][1] = s.y; r.m[2][2] = s.z; return r; }
    // Matrix4x4.LookAt(from, to, up) = TRS(from, LookRotation(to - from, up), 1): columns right, up', forward (Appendix A4)
    static Matrix4x4 LookAt(Vector3 from, Vector3 to, Vector3 up) {
        Vector3 f = Vector3::Normalize(to - from);
        Vector3 r = Vector3::Normalize(Vector3::Cross(up, f));
        Vector3 u = Vector3::Cross(f, r);
        Matrix4x4 o = identity_();
        o.m[0][0] = r.x; o.m[0][1] = r.y; o.m[0][2] = r.z;
        o.m[1][0] = u.x; o.m[1][1] = u.y; o.m[1][2] = u.z;
        o.m[2][0] = f.x; o.m[2][1] = f.y; o.m[2][2] = f.z;
        o.m[3][0] = from.x; o.m[3][1] = from.y; o.m[3][2] = from.z;
        return o;
    }
};
inline float4x4::float4x4(const Matrix4x4& mm)
    : c0(mm.m[0][0], mm.m[0][1], mm.m[0][2], mm.m[0][3]), c1(mm.m[1][0], mm.m[1][1], mm.m[1][2], mm.m[1][3]),
      c2(mm.m[2][0], mm.m[2][1], mm.m[2][2], mm.m[2][3]), c3(mm.m[3][0], mm.m[3][1], mm.m[3][2], mm.m[3][3]) {}

struct Transform {
    Vector3 position;
    Quaternion rotation;
    Vector3 forward;      // rotation * (0,0,1)
    Vector3 up;           // rotation * (0,1,0)
    Vector3 eulerAngles;  // only .x is read by the reference (RenderManager.cs:377): sin(eulerAngles.x) = -forward.y (Appendix A6)
    void set_rotation(Quaternion q) {
        rotation = q;
        forward = q * Vector3(0, 0, 1);
        up = q * Vector3(0, 1, 0);
        // Z-X-Y Euler pitch in degrees, [0, 360)
        float fy = forward.y < -1.0f ? -1.0f : (forward.y > 1.0f ? 1.0f : forward.y);
        float pitch = (float)(std::asin((double)-fy) * (180.0 / 3.14159265358979323846));
        if (pitch < 0.0f) pitch += 360.0f;
        eulerAngles = Vector3(pitch, 0, 0);
    }
};

struct Camera {
    Transform transform;
    float nearClipPlane = 0.05f, farClipPlane = 1000.0f, fieldOfView = 60.0f;
    int pixelWidth = 0, pixelHeight = 0;
    Matrix4x4 worldToCameraMatrix;
    Matrix4x4 nonJitteredProjectionMatrix;
    void RemoveAllCommandBuffers() {}
    template <class E, class C> void AddCommandBuffer(E, C&) {}
    // pose -> matrices (Appendix A2/A3): GL-convention perspective; view = Scale(1,1,-1) * inverse(TRS(pos, rot, 1))
    void update_matrices() {
        float aspect = (float)pixelWidth / (float)pixelHeight;
        float t = (float)std::tan((double)(fieldOfView * Mathf::Deg2Rad * 0.5f));
        float cot = 1.0f / t;
        Matrix4x4 p;
        p.m[0][0] = cot / aspect;
        p.m[1][1] = cot;
        p.m[2][2] = -(farClipPlane + nearClipPlane) / (farClipPlane - nearClipPlane);
        p.m[3][2] = -(2.0f * farClipPlane * nearClipPlane) / (farClipPlane - nearClipPlane);
        p.m[2][3] = -1.0f;
        nonJitteredProjectionMatrix = p;
        Vector3 r = transform.rotation * Vector3(1, 0, 0), u = transform.rotation * Vector3(0, 1, 0), f = transform.rotation * Vector3(0, 0, 1);
        Vector3 pos = transform.position;
        Matrix4x4 v = Matrix4x4::identity_();
        v.m[0][0] = r.x; v.m[1][0] = r.y; v.m[2][0] = r.z; v.m[3][0] = -Vector3::Dot(r, pos);
        v.m[0][1] = u.x; v.m[1][1] = u.y; v.m[2][1] = u.z; v.m[3][1] = -Vector3::Dot(u, pos);
        v.m[0][2] = -f.x; v.m[1][2] = -f.y; v.m[2][2] = -f.z; v.m[3][2] = Vector3::Dot(f, pos);
        worldToCameraMatrix = v;
    }
};

struct Debug {
    template <class... A> static void DrawLine(A&&...) {}
    template <class... A> static void Log(A&&...) {}
};
struct Profiler {
    static void BeginSample(const char*) {}
    static void EndSample() {}
};

// Textures: host memory stand-ins for Texture2D / RenderTexture (handles; copies alias the pixels)
enum class TextureFormat { ARGB32 };
enum class RenderTextureFormat { ARGB32 };
enum class FilterMode { Point, Bilinear };
template <class T>
struct RawTextureData {
    T* p;
    void* GetUnsafePtr() const { return p; }
};
struct Texture2D {
    int width = 0, height = 0;
    FilterMode filterMode = FilterMode::Point;
    std::shared_ptr<std::vector<uint32_t>> own;
    uint32_t* pixels = nullptr;  // borrowed or owned
    Texture2D() {}
    Texture2D(int w, int h, TextureFormat, bool, bool) : width(w), height(h) {
        own = std::make_shared<std::vector<uint32_t>>((size_t)w * h);
        pixels = own->data();
    }
    static Texture2D Borrow(int w, int h, uint32_t* px) { Texture2D t; t.width = w; t.height = h; t.pixels = px; return t; }
    template <class T> RawTextureData<T> GetRawTextureData() const { return RawTextureData<T>{(T*)pixels}; }
    void Apply(bool, bool) {}
};
struct RenderTextureDescriptor {
    int width, height;
    RenderTextureDescriptor(int w, int h, RenderTextureFormat, int, int) : width(w), height(h) {}
};
struct RenderTexture {
    int width = 0, height = 0;
    FilterMode filterMode = FilterMode::Point;
    std::shared_ptr<std::vector<uint32_t>> px;
    RenderTexture() {}
    explicit RenderTexture(const RenderTextureDescriptor& d) : width(d.width), height(d.height), px(std::make_shared<std::vector<uint32_t>>((size_t)d.width * d.height)) {}
};
struct Object {
    template <class T> static void Destroy(T&) {}
};

}  // namespace cpuvox_ref
