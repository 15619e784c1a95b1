// TEST INFRASTRUCTURE ONLY — stand-ins for Unity.Collections 1.2.4, Unity.Jobs, Unity.Burst (the subset the linked reference
// files call). Jobs run inline on Schedule (Parallel.For over batches), which is all the dependency chain of
// RenderManager.DrawSegments needs. Never compiled in this image (no C# toolchain).
using System;
using System.Runtime.InteropServices;
using System.Threading;
using System.Threading.Tasks;

namespace Unity.Burst
{
    public enum FloatPrecision { Standard, High, Medium, Low }
    public enum FloatMode { Default, Strict, Deterministic, Fast }
    [AttributeUsage(AttributeTargets.All)]
    public class BurstCompileAttribute : Attribute
    {
        public BurstCompileAttribute() { }
        public BurstCompileAttribute(FloatPrecision p, FloatMode m) { }
    }
    public struct FunctionPointer<T> where T : Delegate
    {
        readonly T d;
        public FunctionPointer(T d) { this.d = d; }
        public T Invoke => d;
    }
    public static class BurstCompiler
    {
        public static FunctionPointer<T> CompileFunctionPointer<T>(T d) where T : Delegate => new FunctionPointer<T>(d);
    }
}

namespace AOT
{
    [AttributeUsage(AttributeTargets.Method)]
    public class MonoPInvokeCallbackAttribute : Attribute { public MonoPInvokeCallbackAttribute(Type t) { } }
}

namespace Unity.Collections
{
    public enum Allocator { Invalid, None, Temp, TempJob, Persistent }
    public enum NativeArrayOptions { UninitializedMemory, ClearMemory }
    [AttributeUsage(AttributeTargets.Field)] public class ReadOnlyAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.Field)] public class WriteOnlyAttribute : Attribute { }

    public unsafe struct NativeArray<T> : IDisposable where T : struct
    {
        internal void* ptr;
        int length;
        public NativeArray(int length, Allocator allocator, NativeArrayOptions options = NativeArrayOptions.ClearMemory)
        {
            this.length = length;
            long bytes = (long)LowLevel.Unsafe.UnsafeUtility.SizeOf<T>() * Math.Max(length, 1);
            ptr = LowLevel.Unsafe.UnsafeUtility.Malloc(bytes, 16, allocator);
            if (options == NativeArrayOptions.ClearMemory) LowLevel.Unsafe.UnsafeUtility.MemClear(ptr, bytes);
        }
        internal NativeArray(void* borrowed, int length) { ptr = borrowed; this.length = length; }
        public int Length => length;
        public T this[int i]
        {
            get { if ((uint)i >= (uint)length) throw new IndexOutOfRangeException(); return System.Runtime.CompilerServices.Unsafe.Read<T>((byte*)ptr + (long)i * LowLevel.Unsafe.UnsafeUtility.SizeOf<T>()); }
            set { if ((uint)i >= (uint)length) throw new IndexOutOfRangeException(); System.Runtime.CompilerServices.Unsafe.Write((byte*)ptr + (long)i * LowLevel.Unsafe.UnsafeUtility.SizeOf<T>(), value); }
        }
        public void Dispose() { LowLevel.Unsafe.UnsafeUtility.Free(ptr, Allocator.Persistent); ptr = null; length = 0; }
        public static void Copy(NativeArray<T> src, NativeArray<T> dst, int n) => LowLevel.Unsafe.UnsafeUtility.MemCpy(dst.ptr, src.ptr, (long)n * LowLevel.Unsafe.UnsafeUtility.SizeOf<T>());
    }

    public unsafe struct NativeList<T> : IDisposable where T : struct
    {
        // header: [0] length, [1] capacity; data follows in its own allocation
        internal int* header;
        internal void* data;
        public NativeList(int capacity, Allocator allocator)
        {
            header = (int*)LowLevel.Unsafe.UnsafeUtility.Malloc(8, 4, allocator);
            header[0] = 0; header[1] = capacity;
            data = LowLevel.Unsafe.UnsafeUtility.Malloc((long)LowLevel.Unsafe.UnsafeUtility.SizeOf<T>() * Math.Max(capacity, 1), 16, allocator);
        }
        public int Length => Volatile.Read(ref header[0]);
        public T this[int i]
        {
            get { if ((uint)i >= (uint)Length) throw new IndexOutOfRangeException(); return System.Runtime.CompilerServices.Unsafe.Read<T>((byte*)data + (long)i * LowLevel.Unsafe.UnsafeUtility.SizeOf<T>()); }
        }
        public ParallelWriter AsParallelWriter() => new ParallelWriter { header = header, data = data };
        public void Dispose() { LowLevel.Unsafe.UnsafeUtility.Free(data, Allocator.Persistent); LowLevel.Unsafe.UnsafeUtility.Free(header, Allocator.Persistent); header = null; data = null; }
        public unsafe struct ParallelWriter
        {
            internal int* header;
            internal void* data;
            public void AddNoResize(T value)
            {
                int i = Interlocked.Increment(ref header[0]) - 1;
                if (i >= header[1]) throw new InvalidOperationException("NativeList.AddNoResize over capacity");
                System.Runtime.CompilerServices.Unsafe.Write((byte*)data + (long)i * LowLevel.Unsafe.UnsafeUtility.SizeOf<T>(), value);
            }
        }
    }
}

namespace Unity.Collections.LowLevel.Unsafe
{
    [AttributeUsage(AttributeTargets.Field)] public class NativeDisableUnsafePtrRestrictionAttribute : Attribute { }

    public static unsafe class UnsafeUtility
    {
        public static void* Malloc(long size, int alignment, Allocator allocator) => NativeMemory.AlignedAlloc((nuint)Math.Max(size, 1), (nuint)Math.Max(alignment, 8));
        public static void Free(void* p, Allocator allocator) { if (p != null) NativeMemory.AlignedFree(p); }
        public static void MemClear(void* p, long size) => NativeMemory.Clear(p, (nuint)size);
        public static void MemCpy(void* dst, void* src, long size) => Buffer.MemoryCopy(src, dst, size, size);
        public static int SizeOf<T>() where T : struct => System.Runtime.CompilerServices.Unsafe.SizeOf<T>();
        public static int AlignOf<T>() where T : struct => Math.Min(SizeOf<T>(), 8) < 4 ? 4 : 4;
        public static void CopyStructureToPtr<T>(ref T s, void* p) where T : struct => System.Runtime.CompilerServices.Unsafe.Write(p, s);
        public static void CopyPtrToStructure<T>(void* p, out T s) where T : struct => s = System.Runtime.CompilerServices.Unsafe.Read<T>(p);
    }

    public static unsafe class NativeArrayUnsafeUtility
    {
        public static void* GetUnsafePtr<T>(this NativeArray<T> a) where T : struct => a.ptr;
        public static void* GetUnsafeReadOnlyPtr<T>(this NativeArray<T> a) where T : struct => a.ptr;
        public static NativeArray<T> ConvertExistingDataToNativeArray<T>(void* p, int length, Allocator a) where T : struct => new NativeArray<T>(p, length);
    }
}

namespace Unity.Jobs
{
    public struct JobHandle { public void Complete() { } }
    public interface IJobParallelFor { void Execute(int index); }

    public static class IJobParallelForExtensions
    {
        public static int Threads = Environment.ProcessorCount;
        // Runs to completion before returning; batches are handed out like Unity's work stealing (batch size 1 for RenderJob).
        public static JobHandle Schedule<T>(this T job, int arrayLength, int innerloopBatchCount, JobHandle dependsOn = default) where T : struct, IJobParallelFor
        {
            if (Threads <= 1 || arrayLength <= innerloopBatchCount)
            {
                for (int i = 0; i < arrayLength; i++) job.Execute(i);
                return default;
            }
            int batches = (arrayLength + innerloopBatchCount - 1) / innerloopBatchCount;
            Parallel.For(0, batches, new ParallelOptions { MaxDegreeOfParallelism = Threads }, b =>
            {
                T local = job;
                int end = Math.Min(arrayLength, (b + 1) * innerloopBatchCount);
                for (int i = b * innerloopBatchCount; i < end; i++) local.Execute(i);
            });
            return default;
        }
    }
}
