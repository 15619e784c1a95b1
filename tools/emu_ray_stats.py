"""Developer tool (CPU): per-ray event counts of the Phase-1 kernel through the SIMT emulator built with -DCVX_EMU_STATS
(make -C tools/simt_emu EXTRA=-DCVX_EMU_STATS OUT=libcvx_emu_stats.so): batches, column entries, re-narrowings, rounds, commit candidates,
commits, round occupancy.  python tools/emu_ray_stats.py mill 59 | terrain"""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import cpuvox_b200 as cv
L = C.CDLL('/root/repo/tools/simt_emu/libcvx_emu_stats.so')
L.emu_world_create.restype = C.c_void_p
L.emu_world_create.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
L.emu_phase1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5
L.emu_set_stats.argtypes=[C.c_void_p]
kind = sys.argv[1]
if kind == 'mill':
    world = cv.World.from_obj('tests/data/mill.obj', 1024); W,H=3840,2160
    poses = cv.benchmark_path(world.dims, 60, far_clip=2.0*world.max_dimension); pose = poses[int(sys.argv[2])]
else:
    world = cv.World.synthetic(0,(2048,2048,2048),seed=1234); W,H=1920,1080
    pose = cv.CameraPose.from_euler((1024.0,1700.0,1024.0),(60.0,30.0,0.0),far_clip=4096.0)
lods = cv.setup_lods(world.max_dimension, W, H)
s = cv.frame_setup(pose, W, H, lods, world.dims[1])
n=len(world.blobs); blobs=[np.ascontiguousarray(b) for b in world.blobs]
dims=(C.c_int32*3)(*world.dims); ptrs=(C.c_void_p*n)(*[b.ctypes.data for b in blobs]); sizes=(C.c_int64*n)(*[b.nbytes for b in blobs]); counts=(C.c_int32*n)(*world.column_counts)
w = L.emu_world_create(n,dims,ptrs,sizes,counts)
total = sum(max(0,s.segments[k].ray_count) for k in range(4))
stats = np.zeros((total,16),dtype=np.uint64); L.emu_set_stats(stats.ctypes.data)
td=np.zeros((W+2*H,H),dtype=np.uint32); lr=np.zeros((2*W+H,W),dtype=np.uint32)
step = max(1,total//96)
for r0 in range(0,total,step):
    L.emu_phase1(w, C.addressof(s), W, H, td.ctypes.data, lr.ctypes.data, None, 1, 32, 1, r0, r0+1)
st = stats[::step].astype(np.float64)
names=["batches","select","renarrow","col passes","rounds","cand","commits","caps","nearplane","lanes","cols cached","cols considered"]
print("rays sampled",len(st))
for i,nm in enumerate(names): print(f"{nm:16s} mean {st[:,i].mean():10.1f}  max {st[:,i].max():10.0f}")
print("lanes/round %.1f cols cached/round %.2f considered/round %.2f  col passes/round %.2f" % (st[:,9].sum()/st[:,4].sum(), st[:,10].sum()/st[:,4].sum(), st[:,11].sum()/st[:,4].sum(), st[:,3].sum()/st[:,4].sum()))
