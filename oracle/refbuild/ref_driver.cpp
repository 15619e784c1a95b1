// ref_driver.cpp — TEST INFRASTRUCTURE ONLY. C ABI over the reference's own code as translated by cs2cpp.py (ref_gen.hpp,
// generated at build time from /root/reference, never committed). Built into oracle/_ref/libcpuvox_ref.so.
//
// Everything that computes here is the reference's text: World / RLEColumn / WorldAllocator (World.cs), SegmentDDAData,
// CameraData, RayBuffer.Native, RenderManager.DrawWorld / DrawSegments / BlitSegments / GetGenericSegmentParameters /
// vanishing point, all four jobs of DrawSegmentRayJob.cs, WorldBuilder.RLEColumnBuilder + World.DownSample,
// VoxelizerHelper.GetVoxelsInternal, SimpleMesh.Remap_Internal, and the default variant of RayBufferBlit.shader's frag.
// This file only marshals plain C structs in and out, and restates the few lines of glue that live in Unity MonoBehaviour
// code the translation does not cover (cited where they appear).
#include <cstdint>
#include <cstring>
#include <mutex>

#include "ref_gen.hpp"

namespace cpuvox_ref {
int g_job_threads = 1;
Vector3 Vector3::zero(0, 0, 0);
Vector3 Vector3::one(1, 1, 1);
Vector3 Vector3::up(0, 1, 0);
Vector3 Vector3::forward(0, 0, 1);
Color Color::red(1, 0, 0, 1);
Matrix4x4 Matrix4x4::identity = Matrix4x4::identity_();
}  // namespace cpuvox_ref

using namespace cpuvox_ref;

extern "C" {

#define REF_LOD_LEVELS 6
static_assert(REF_LOD_LEVELS == UnityManager::LOD_LEVELS, "LOD_LEVELS of the reference changed");

// layouts equal oracle/cpuvox_oracle.h and include/cpuvox_b200.h (tests memcpy between them)
typedef struct ref_segment {
    float min_screen[2], max_screen[2], cam_local_plane_ray_min[2], cam_local_plane_ray_max[2];
    int32_t ray_count;
} ref_segment;
typedef struct ref_camera {
    float world_to_screen[16];
    float position_xz[2];
    float position_y;
    int32_t inverse_element_iteration_direction;
    float far_clip;
    float lod_distances[REF_LOD_LEVELS];
} ref_camera;
typedef struct ref_frame_setup {
    ref_segment segments[4];
    ref_camera camera;
    float vanishing_point_screen[2];
} ref_frame_setup;
typedef struct ref_pose {
    float position[3];
    float rotation[4];
    float fov_y_degrees, near_clip, far_clip;
    int32_t pixel_width, pixel_height;
} ref_pose;

struct ref_world {
    ManagedArray<World> lods{REF_LOD_LEVELS};
    int3 dims;
};

static thread_local std::string g_err;
const char* ref_last_error(void) { return g_err.c_str(); }
#define REF_TRY try {
#define REF_CATCH(rc)                      \
    }                                      \
    catch (const std::exception& e) {      \
        g_err = e.what();                  \
        return rc;                         \
    }

ref_world* ref_world_create(int32_t dx, int32_t dy, int32_t dz) {
    ref_world* w = new ref_world();
    w->dims = int3(dx, dy, dz);
    return w;
}
// blob: the reference allocator's memory (World.cs:273-293); borrowed. Goes through `new World(dimensions, lod, data)`
// (World.cs:37-44), so the header count is World.ColumnCount (World.cs:17) — returned for the caller to compare.
int ref_world_set_lod(ref_world* w, int32_t lod, const void* blob, int64_t bytes, int32_t column_count) {
    REF_TRY
    if (!w || lod < 0 || lod >= REF_LOD_LEVELS || !blob) return -1;
    World wl(w->dims, lod, const_cast<void*>(blob));
    if (wl.ColumnCount() != column_count) { g_err = "ColumnCount differs from World.cs:17"; return -2; }
    if ((int64_t)wl.ColumnCount() * 12 > bytes) { g_err = "blob shorter than its headers"; return -3; }
    w->lods[lod] = wl;
    return 0;
    REF_CATCH(-9)
}
int32_t ref_world_column_count(int32_t dx, int32_t dy, int32_t dz, int32_t lod) {
    World wl(int3(dx, dy, dz), lod, (void*)nullptr);
    return wl.ColumnCount();
}
void ref_world_free(ref_world* w) { delete w; }

static void camera_from_pose(const ref_pose* p, Camera& cam) {
    cam.transform.position = Vector3(p->position[0], p->position[1], p->position[2]);
    cam.transform.set_rotation(Quaternion(p->rotation[0], p->rotation[1], p->rotation[2], p->rotation[3]));
    cam.fieldOfView = p->fov_y_degrees;
    cam.nearClipPlane = p->near_clip;
    cam.farClipPlane = p->far_clip;
    cam.pixelWidth = p->pixel_width;
    cam.pixelHeight = p->pixel_height;
    cam.update_matrices();
}

static void export_setup(const NativeArray<RenderManager_SegmentData>& segments, CameraData& cd, float2 vp, ref_frame_setup* out) {
    memset(out, 0, sizeof *out);
    for (int k = 0; k < 4; k++) {
        const RenderManager_SegmentData& s = segments[k];
        ref_segment& o = out->segments[k];
        o.min_screen[0] = s.MinScreen.x; o.min_screen[1] = s.MinScreen.y;
        o.max_screen[0] = s.MaxScreen.x; o.max_screen[1] = s.MaxScreen.y;
        o.cam_local_plane_ray_min[0] = s.CamLocalPlaneRayMin.x; o.cam_local_plane_ray_min[1] = s.CamLocalPlaneRayMin.y;
        o.cam_local_plane_ray_max[0] = s.CamLocalPlaneRayMax.x; o.cam_local_plane_ray_max[1] = s.CamLocalPlaneRayMax.y;
        o.ray_count = s.RayCount;
    }
    const float4* c[4] = {&cd.WorldToScreenMatrix.c0, &cd.WorldToScreenMatrix.c1, &cd.WorldToScreenMatrix.c2, &cd.WorldToScreenMatrix.c3};
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) out->camera.world_to_screen[j * 4 + i] = (*c[j])[i];
    out->camera.position_xz[0] = cd.PositionXZ.x;
    out->camera.position_xz[1] = cd.PositionXZ.y;
    out->camera.position_y = cd.PositionY;
    out->camera.inverse_element_iteration_direction = cd.InverseElementIterationDirection ? 1 : 0;
    out->camera.far_clip = cd.FarClip;
    for (int i = 0; i < REF_LOD_LEVELS; i++) out->camera.lod_distances[i] = cd.LODDistances[i];
    out->vanishing_point_screen[0] = vp.x;
    out->vanishing_point_screen[1] = vp.y;
}

// a1-a5: the head of RenderManager.DrawWorld (RenderManager.cs:119-152) — vanishing point, the four guarded
// GetGenericSegmentParameters calls, the CameraData constructor. The guards are restated here (four `if`s); the functions
// they call are the reference's. ref_draw_world below runs the reference's own DrawWorld, and a test requires both to agree.
int ref_frame_setup_from_pose(const ref_pose* pose, const float lod_distances[REF_LOD_LEVELS], int32_t world_dim_y, ref_frame_setup* out) {
    REF_TRY
    Camera camera;
    camera_from_pose(pose, camera);
    int screenWidth = pose->pixel_width, screenHeight = pose->pixel_height;
    float3 vanishingPointWorldSpace = RenderManager::CalculateVanishingPointWorld(camera);
    float2 vanishingPointScreenSpace = RenderManager::ProjectVanishingPointScreenToWorld(camera, vanishingPointWorldSpace);
    float2 screen = float2(screenWidth, screenHeight);
    NativeArray<RenderManager_SegmentData> segments(4, Allocator::Temp, NativeArrayOptions::ClearMemory);
    if (vanishingPointScreenSpace.y < screenHeight)
        segments[0] = RenderManager::GetGenericSegmentParameters(camera, screen, vanishingPointScreenSpace, screenHeight - vanishingPointScreenSpace.y, float2(0, 1), 1, world_dim_y);
    if (vanishingPointScreenSpace.y > 0.f)
        segments[1] = RenderManager::GetGenericSegmentParameters(camera, screen, vanishingPointScreenSpace, vanishingPointScreenSpace.y, float2(0, -1), 1, world_dim_y);
    if (vanishingPointScreenSpace.x < screenWidth)
        segments[2] = RenderManager::GetGenericSegmentParameters(camera, screen, vanishingPointScreenSpace, screenWidth - vanishingPointScreenSpace.x, float2(1, 0), 0, world_dim_y);
    if (vanishingPointScreenSpace.x > 0.f)
        segments[3] = RenderManager::GetGenericSegmentParameters(camera, screen, vanishingPointScreenSpace, vanishingPointScreenSpace.x, float2(-1, 0), 0, world_dim_y);
    ManagedArray<float> lods(REF_LOD_LEVELS);
    for (int i = 0; i < REF_LOD_LEVELS; i++) lods[i] = lod_distances[i];
    CameraData camData(camera, lods, screen);
    export_setup(segments, camData, vanishingPointScreenSpace, out);
    segments.Dispose();
    return 0;
    REF_CATCH(-9)
}

static void import_setup(const ref_frame_setup* s, NativeArray<RenderManager_SegmentData>& segments, CameraData& cd, float2& vp) {
    for (int k = 0; k < 4; k++) {
        RenderManager_SegmentData& o = segments[k];
        const ref_segment& i = s->segments[k];
        o.MinScreen = float2(i.min_screen[0], i.min_screen[1]);
        o.MaxScreen = float2(i.max_screen[0], i.max_screen[1]);
        o.CamLocalPlaneRayMin = float2(i.cam_local_plane_ray_min[0], i.cam_local_plane_ray_min[1]);
        o.CamLocalPlaneRayMax = float2(i.cam_local_plane_ray_max[0], i.cam_local_plane_ray_max[1]);
        o.RayCount = i.ray_count;
    }
    float4* c[4] = {&cd.WorldToScreenMatrix.c0, &cd.WorldToScreenMatrix.c1, &cd.WorldToScreenMatrix.c2, &cd.WorldToScreenMatrix.c3};
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) (*c[j])[i] = s->camera.world_to_screen[j * 4 + i];
    cd.PositionXZ = float2(s->camera.position_xz[0], s->camera.position_xz[1]);
    cd.PositionY = s->camera.position_y;
    cd.InverseElementIterationDirection = s->camera.inverse_element_iteration_direction != 0;
    cd.FarClip = s->camera.far_clip;
    for (int i = 0; i < REF_LOD_LEVELS; i++) cd.LODDistances[i] = s->camera.lod_distances[i];
    vp = float2(s->vanishing_point_screen[0], s->vanishing_point_screen[1]);
}

// raybuffers as the reference keeps them: partial textures of 256 rays (RayBuffer.cs:18-45), here borrowed slices of the
// caller's flat buffer (row r of the flat buffer = partial r>>8, row r&255 — exactly Native.GetRayColumn's addressing).
static void borrow_partials(RayBuffer& rb, uint32_t* flat, int row_len, int rows) {
    int unused = 0;
    int n = rb.GetPartialsCount(rows, unused);
    rb.Partials = ManagedArray<Texture2D>(n);
    for (int i = 0; i < n; i++) rb.Partials[i] = Texture2D::Borrow(row_len, RayBuffer::RAYS_PER_PARTIAL, flat + (size_t)i * RayBuffer::RAYS_PER_PARTIAL * row_len);
}

// a6-a17: RenderManager.DrawSegments (RenderManager.cs:258-372) on caller memory. td: (W+2H) rows of H pixels, lr: (2W+H) rows
// of W pixels; both must be padded to a multiple of 256 rows (the reference's partial textures are 256 rays tall).
int ref_render_raybuffers(const ref_world* w, const ref_frame_setup* setup, int32_t W, int32_t H, uint32_t* td, uint32_t* lr, int32_t n_threads) {
    REF_TRY
    if (!w || !setup || !td || !lr) return -1;
    g_job_threads = n_threads <= 0 ? (int)std::thread::hardware_concurrency() : n_threads;
    NativeArray<RenderManager_SegmentData> segments(4, Allocator::Temp, NativeArrayOptions::ClearMemory);
    CameraData camData;
    float2 vp;
    import_setup(setup, segments, camData, vp);
    RayBuffer tdManaged, lrManaged;
    borrow_partials(tdManaged, td, H, W + 2 * H);
    borrow_partials(lrManaged, lr, W, 2 * W + H);
    RayBuffer_Native tdNative = tdManaged.GetNativeData(Allocator::TempJob);
    RayBuffer_Native lrNative = lrManaged.GetNativeData(Allocator::TempJob);
    RenderManager::DrawSegments(segments, w->lods.data(), camData, W, H, vp, tdNative, lrNative, tdManaged, lrManaged);
    tdNative.Dispose();
    lrNative.Dispose();
    segments.Dispose();
    return 0;
    REF_CATCH(-9)
}

// a18: RenderManager.BlitSegments (RenderManager.cs:199-256) + RayBufferBlit.shader frag, through the stand-in rasteriser.
// td / lr: flat raybuffers of exactly (W+2H) x H and (2W+H) x W pixels; frame: W x H, row 0 = bottom.
int ref_blit(const ref_frame_setup* setup, int32_t W, int32_t H, const uint32_t* td, const uint32_t* lr, uint32_t* frame) {
    REF_TRY
    NativeArray<RenderManager_SegmentData> segments(4, Allocator::Temp, NativeArrayOptions::ClearMemory);
    CameraData camData;
    float2 vp;
    import_setup(setup, segments, camData, vp);
    RenderTexture tdTex(RenderTextureDescriptor(H, W + 2 * H, RenderTextureFormat::ARGB32, 0, 0));
    RenderTexture lrTex(RenderTextureDescriptor(W, 2 * W + H, RenderTextureFormat::ARGB32, 0, 0));
    memcpy(tdTex.px->data(), td, tdTex.px->size() * 4);
    memcpy(lrTex.px->data(), lr, lrTex.px->size() * 4);
    Camera camera;
    Material material;
    Mesh mesh;
    CommandBuffer commands;
    commands.target = frame;
    commands.targetW = W;
    commands.targetH = H;
    RenderManager::BlitSegments(camera, material, mesh, tdTex, lrTex, segments, vp, float2(W, H), commands);
    segments.Dispose();
    return 0;
    REF_CATCH(-9)
}

// The reference's whole frame: `new RenderManager()` + `DrawWorld` (RenderManager.cs:25-41,111-194), exactly as
// UnityManager.LateUpdate drives it (UnityManager.cs:166-187; LimitRotationHorizon is applied by the caller to the pose).
// Outputs: flat raybuffers (rows as in RayBuffer.Native.GetRayColumn), the frame, and nothing else.
int ref_draw_world(const ref_world* w, const ref_pose* pose, const float lod_distances[REF_LOD_LEVELS], uint32_t* td, uint32_t* lr, uint32_t* frame, int32_t n_threads) {
    REF_TRY
    static std::mutex mu;  // Screen.width/height are process globals, like Unity's
    std::lock_guard<std::mutex> lock(mu);
    g_job_threads = n_threads <= 0 ? (int)std::thread::hardware_concurrency() : n_threads;
    const int W = pose->pixel_width, H = pose->pixel_height;
    Screen::width = W;
    Screen::height = H;
    RenderManager rm;
    Camera camera;
    camera_from_pose(pose, camera);
    ManagedArray<float> lods(REF_LOD_LEVELS);
    for (int i = 0; i < REF_LOD_LEVELS; i++) lods[i] = lod_distances[i];
    Material material;
    rm.commandBuffer.target = frame;
    rm.commandBuffer.targetW = W;
    rm.commandBuffer.targetH = H;
    rm.DrawWorld(material, w->lods, camera, camera, lods);
    RayBuffer& tdb = rm.rayBufferTopDown[rm.bufferIndex];
    RayBuffer& lrb = rm.rayBufferLeftRight[rm.bufferIndex];
    if (td) memcpy(td, tdb.FinalTexture.px->data(), (size_t)(W + 2 * H) * H * 4);
    if (lr) memcpy(lr, lrb.FinalTexture.px->data(), (size_t)(2 * W + H) * W * 4);
    rm.Destroy();
    return 0;
    REF_CATCH(-9)
}

// Test hook: cell sequence of one ray as ExecuteRay's loop head walks it (DrawSegmentRayJob.cs:235-243,613): SegmentDDAData only.
int ref_dda_walk(const float start[2], const float dir[2], const float lod_distances[REF_LOD_LEVELS], float far_clip, int32_t max_steps,
                 int32_t* out_cells, float* out_dists) {
    SegmentDDAData ray(float2(start[0], start[1]), float2(dir[0], dir[1]));
    int lod = 0, voxelScale = 1, n = 0;
    float lodMax = lod_distances[0];
    while (n < max_steps) {
        if (ray.IntersectionDistances().x >= lodMax) {
            ray.NextLOD(voxelScale);
            lod++;
            voxelScale *= 2;
            lodMax = lod_distances[lod];
        }
        out_cells[n * 3 + 0] = ray.Position().x;
        out_cells[n * 3 + 1] = ray.Position().y;
        out_cells[n * 3 + 2] = lod;
        out_dists[n * 2 + 0] = ray.IntersectionDistances().x;
        out_dists[n * 2 + 1] = ray.IntersectionDistances().y;
        n++;
        if (ray.Step(far_clip)) break;
    }
    return n;
}

// f2: world production with the reference's own builder code. Mesh in, LOD blobs out.
//   positions: n x {x, y, z} floats, colors32: n x {r, g, b, a} bytes (SimpleMesh.Vertex.Color is a Color32: ObjModel.cs:136
//   stores the parsed float colour through Unity's implicit Color -> Color32 rounding), indices: triangles.
// Steps = UnityManager.cs:340-366: SimpleMesh.Rescale (Remap_Internal), WorldBuilder.Import (its per-triangle body restated
// here single-threaded: VoxelizerHelper.GetVoxels + RLEColumnBuilder.SetVoxel, WordBuilder.cs:71-91, no materials),
// ToLOD0World, DownSample(j) for j = 1..5. The blobs are the allocators' memory (GetStartPointer / GetByteLength).
struct ref_built_world {
    ManagedArray<World> lods{REF_LOD_LEVELS};
    int3 dims;
    int voxels[REF_LOD_LEVELS];
};
ref_built_world* ref_build_world_from_mesh(const float* positions, const uint8_t* colors32, int32_t n_vertices, const int32_t* indices, int32_t n_indices,
                                           int32_t max_dimension, int32_t flip_x, int32_t flip_y, int32_t flip_z, int32_t n_lods) {
    try {
        std::vector<SimpleMesh_Vertex> verts((size_t)n_vertices);
        for (int i = 0; i < n_vertices; i++) {
            const float* v = positions + (size_t)i * 3;
            const uint8_t* c = colors32 + (size_t)i * 4;
            verts[i].Position() = float3(v[0], v[1], v[2]);
            verts[i].Color = Color32(c[0], c[1], c[2], c[3]);
            verts[i].MaterialIndex = -1;
        }
        std::vector<int> idx(indices, indices + n_indices);
        float3 flips(flip_x ? -1.f : 1.f, flip_y ? -1.f : 1.f, flip_z ? -1.f : 1.f);
        int3 worldDimensions;
        SimpleMesh::Remap_Internal(verts.data(), n_vertices, (float)max_dimension, flips, worldDimensions);
        WorldBuilder builder(worldDimensions.x, worldDimensions.y, worldDimensions.z);
        {
            VoxelizerHelper_GetVoxelsContext context;
            context.maxDimensions = builder.Dimensions() - 1;
            std::vector<VoxelizerHelper_VoxelizedPosition> buf((size_t)WorldBuilder::VOXELIZE_BUFFER_MAX);
            context.positions = buf.data();
            context.positionLength = WorldBuilder::VOXELIZE_BUFFER_MAX;
            context.verts = verts.data();
            context.indices = idx.data();
            for (int i = 0; i + 2 < n_indices; i += 3) {
                VoxelizerHelper::GetVoxelsInternal(context, i);
                int written = context.writtenVoxelCount;
                for (int j = 0; j < written; j++) {
                    VoxelizerHelper_VoxelizedPosition pos = context.positions[j];
                    WorldBuilder_RLEColumnBuilder& column = builder.WorldColumns[pos.XZIndex];
                    column.SetVoxel(pos.Y, pos.Color);
                }
            }
        }
        ref_built_world* out = new ref_built_world();
        out->dims = worldDimensions;
        memset(out->voxels, 0, sizeof out->voxels);
        out->lods[0] = builder.ToLOD0World(out->voxels[0]);
        for (int j = 1; j < n_lods && j < REF_LOD_LEVELS; j++) out->lods[j] = out->lods[0].DownSample(j, out->voxels[j]);
        return out;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_built_world_info(const ref_built_world* b, int32_t dims[3], int64_t bytes[REF_LOD_LEVELS], int32_t column_counts[REF_LOD_LEVELS], int32_t voxels[REF_LOD_LEVELS]) {
    ref_built_world* bw = const_cast<ref_built_world*>(b);
    dims[0] = bw->dims.x; dims[1] = bw->dims.y; dims[2] = bw->dims.z;
    for (int j = 0; j < REF_LOD_LEVELS; j++) {
        bool ex = bw->lods[j].Exists();
        bytes[j] = ex ? bw->lods[j].Storage.GetByteLength() : 0;
        column_counts[j] = ex ? bw->lods[j].ColumnCount() : 0;
        voxels[j] = bw->voxels[j];
    }
}
int ref_built_world_copy_blob(const ref_built_world* b, int32_t lod, void* dst, int64_t bytes) {
    ref_built_world* bw = const_cast<ref_built_world*>(b);
    if (lod < 0 || lod >= REF_LOD_LEVELS || !bw->lods[lod].Exists()) return -1;
    if (bytes != bw->lods[lod].Storage.GetByteLength()) return -2;
    memcpy(dst, bw->lods[lod].Storage.GetStartPointer(), (size_t)bytes);
    return 0;
}
// f1: WorldSaveFile.Serialize (WorldSaveFile.cs:8-55) of the built LODs — a .world file written by the reference's own code
int ref_built_world_save(const ref_built_world* b, const char* path) {
    REF_TRY
    ref_built_world* bw = const_cast<ref_built_world*>(b);
    for (int j = 0; j < REF_LOD_LEVELS; j++)
        if (!bw->lods[j].Exists()) { g_err = "Serialize needs all LODs (UnityManager.cs:358-366 builds six)"; return -1; }
    WorldSaveFile::Serialize(bw->lods, string(path));
    return 0;
    REF_CATCH(-9)
}
// WorldSaveFile.Deserialize (WorldSaveFile.cs:57-93): the worlds it returns, ready for ref_render_raybuffers / ref_draw_world
ref_world* ref_world_load(const char* path, int32_t out_dims[3], int32_t* out_world_count) {
    try {
        ManagedArray<World> worlds = WorldSaveFile::Deserialize(string(path));
        ref_world* w = new ref_world();
        w->dims = worlds[0].Dimensions();
        for (int j = 0; j < worlds.Length && j < REF_LOD_LEVELS; j++) w->lods[j] = worlds[j];
        out_dims[0] = w->dims.x; out_dims[1] = w->dims.y; out_dims[2] = w->dims.z;
        *out_world_count = worlds.Length;
        return w;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_built_world_free(ref_built_world* b) {
    if (!b) return;
    for (int j = 0; j < REF_LOD_LEVELS; j++)
        if (b->lods[j].Exists()) b->lods[j].Dispose();
    delete b;
}

int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
const char* ref_describe(void) {
    return "pipliz/cpuvox C# sources translated to C++ by oracle/refbuild/cs2cpp.py, g++ -O2 -ffp-contract=off (IEEE fp32, no Burst)";
}

}  // extern "C"
