#!/bin/bash
# loader with the host-side mask front (route auto / device) and the L2 prefetch-ahead distance of the c2 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
line() { python - "$1" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["config"]
print("   value %.4g  ms/pass %s ms/volume %s e2e %.4g  frac %.3f" % (d["value"], c.get("ms_per_pass"), c.get("ms_per_volume"), d["e2e"]["value"], d["roofline"]["frac"]))
P
}
{
nproc
echo "== loader tests"
timeout 900 python -m pytest tests/test_gpu_loader_roi.py tests/test_host_loader_front.py -q -x 2>&1 | tail -3
echo "== host front, 160^3 x 6"
T2FIT_HOST_PROFILE=1 python - <<'P' 2>&1 | tail -6
import numpy as np, ctypes as C, time
from fetal_t2mapping_b200 import _abi
lib=_abi.load_library()
rng=np.random.default_rng(0)
n=160**3
z,y,x=np.mgrid[:160,:160,:160]
m=(((z-80)/45.)**2+((y-80)/50.)**2+((x-80)/40.)**2<1)
masks=[m.astype(np.uint8) for _ in range(6)]
planes=[rng.random(m.shape,dtype=np.float32) for _ in range(6)]
mo=np.empty(n,np.uint8); idx=np.empty(n,np.int64); cnt=C.c_int64(); SOA=np.zeros(6*n,np.float32)
best=[1e9,1e9]
for rep in range(6):
    t0=time.perf_counter()
    lib.t2fit_host_mask_union_indices((C.c_void_p*6)(*[a.ctypes.data for a in masks]),6,0,None,0,n,mo.ctypes.data,idx.ctypes.data,C.byref(cnt))
    t1=time.perf_counter(); M=cnt.value
    lib.t2fit_host_gather_planes((C.c_void_p*6)(*[a.ctypes.data for a in planes]),6,4,idx.ctypes.data,M,n,SOA.ctypes.data,M)
    t2=time.perf_counter()
    best=[min(best[0],(t1-t0)*1e3),min(best[1],(t2-t1)*1e3)]
print("union+idx ms %.3f gather ms %.3f M %d"%(best[0],best[1],M))
P
for r in auto device auto device; do
echo "== bench c4 route $r"
T2FIT_BENCH_C4_ROUTE=$r timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n_c4_$r.json 2> gpurun_out/n_c4_$r.err; line gpurun_out/n_c4_$r.json
done
for rep in 1 2; do for a in 0 740 1480 2960 5920; do
echo "== bench c2 ahead $a (rep $rep)"
T2FIT_PREFETCH_AHEAD=$a timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/n_c2_a${a}_$rep.json 2> gpurun_out/n_c2_a${a}_$rep.err; line gpurun_out/n_c2_a${a}_$rep.json
done; done
} 2>&1 | tee gpurun_out/n_job.log
