// TEST INFRASTRUCTURE ONLY — runs the reference's RenderManager.DrawWorld (linked by path from the reference checkout) on a case
// file written by tools/dotnet_ref.py and dumps both raybuffers; `convert` runs the reference's .obj -> .world pipeline
// (UnityManager.cs:340-366) so that a .world written by the reference's own WorldSaveFile.Serialize exists for the f1 interop test.
//   cpuvox_ref_dotnet render <case.bin> <out.bin> [threads]
//   cpuvox_ref_dotnet convert <model.obj> <out.world> <maxDimension>
// Case file (little endian): int32 W, H, dims[3], nLods; per LOD: int32 columnCount, int64 bytes, blob; then
// float position[3], rotation[4], fov, near, far, lodDistances[6].  Output: TD (W+2H)xH then LR (2W+H)xW uint32 pixels.
// Never compiled in this image (no C# toolchain); see oracle/dotnet/CpuvoxRef.csproj.
using System;
using System.IO;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Jobs;
using Unity.Mathematics;
using UnityEngine;

public static unsafe class Driver
{
    public static int Main(string[] args)
    {
        if (args.Length >= 3 && args[0] == "render") return Render(args[1], args[2], args.Length > 3 ? int.Parse(args[3]) : Environment.ProcessorCount);
        if (args.Length >= 4 && args[0] == "convert") return Convert(args[1], args[2], int.Parse(args[3]));
        Console.Error.WriteLine("usage: render <case> <out> [threads] | convert <obj> <world> <maxDimension>");
        return 2;
    }

    static int Render(string casePath, string outPath, int threads)
    {
        IJobParallelForExtensions.Threads = threads;
        using var f = new BinaryReader(File.OpenRead(casePath));
        int W = f.ReadInt32(), H = f.ReadInt32();
        int3 dims = new int3(f.ReadInt32(), f.ReadInt32(), f.ReadInt32());
        int nLods = f.ReadInt32();
        World[] worlds = new World[UnityManager.LOD_LEVELS];
        for (int lod = 0; lod < nLods; lod++)
        {
            int columnCount = f.ReadInt32();
            long bytes = f.ReadInt64();
            void* mem = UnsafeUtility.Malloc(bytes, 16, Allocator.Persistent);
            var span = new Span<byte>(mem, checked((int)bytes));
            f.BaseStream.ReadExactly(span);
            worlds[lod] = new World(dims, lod, mem);
            if (worlds[lod].ColumnCount != columnCount) throw new InvalidDataException($"ColumnCount {worlds[lod].ColumnCount} != {columnCount} at LOD {lod}");
        }
        Camera cam = new Camera();
        cam.transform.position = new Vector3(f.ReadSingle(), f.ReadSingle(), f.ReadSingle());
        cam.transform.rotation = new Quaternion(f.ReadSingle(), f.ReadSingle(), f.ReadSingle(), f.ReadSingle());
        cam.fieldOfView = f.ReadSingle(); cam.nearClipPlane = f.ReadSingle(); cam.farClipPlane = f.ReadSingle();
        cam.pixelWidth = W; cam.pixelHeight = H;
        float[] lods = new float[UnityManager.LOD_LEVELS];
        for (int i = 0; i < lods.Length; i++) lods[i] = f.ReadSingle();

        Screen.width = W; Screen.height = H;
        RenderManager rm = new RenderManager();
        rm.DrawWorld(new Material(), worlds, cam, cam, lods);

        // the active buffer set is index 0 (bufferIndex starts at 0 and SwapBuffers was not called)
        var flags = System.Reflection.BindingFlags.NonPublic | System.Reflection.BindingFlags.Instance;
        RayBuffer td = ((RayBuffer[])typeof(RenderManager).GetField("rayBufferTopDown", flags).GetValue(rm))[0];
        RayBuffer lr = ((RayBuffer[])typeof(RenderManager).GetField("rayBufferLeftRight", flags).GetValue(rm))[0];
        using var o = File.Create(outPath);
        o.Write(new ReadOnlySpan<byte>(td.FinalTexture.pixels, (W + 2 * H) * H * 4));
        o.Write(new ReadOnlySpan<byte>(lr.FinalTexture.pixels, (2 * W + H) * W * 4));
        return 0;
    }

    static int Convert(string objPath, string worldPath, int maxDimension)
    {
        SimpleMesh mesh = ObjModel.Import(objPath, false);
        int3 worldDimensions = mesh.Rescale(maxDimension, new float3(-1f, 1f, 1f));   // UI defaults: flip X only (UnityManager.cs:27)
        WorldBuilder builder = new WorldBuilder(worldDimensions.x, worldDimensions.y, worldDimensions.z);
        builder.Import(mesh);
        mesh.Dispose();
        World[] worldLODs = new World[UnityManager.LOD_LEVELS];
        worldLODs[0] = builder.ToLOD0World(out int lod0VoxelCount);
        for (int j = 1; j < UnityManager.LOD_LEVELS; j++) worldLODs[j] = worldLODs[0].DownSample(j, out int voxelCount);
        WorldSaveFile.Serialize(worldLODs, worldPath);
        Console.WriteLine($"{worldDimensions} {lod0VoxelCount} voxels -> {worldPath}");
        return 0;
    }
}
