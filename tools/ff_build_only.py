"""GPU box: build the form-factor matrix of one bench workload and nothing else (the target of the ncu captures of k_ff_tiles)."""
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
import daisyriot_b200 as dz  # noqa: E402

name = sys.argv[1]
sc, wl, E, M, tmp = bench.make_workload(name)
p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=dz.msvc_sample_pattern(1))
p.cudaCalculateRadiosityMatrix()
print(name, p.stats())
p.close()
