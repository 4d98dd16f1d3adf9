import sys, numpy as np
sys.path.insert(0,'.')
import daisyriot_b200 as dz
from daisyriot_b200 import scenes
from oracle import pyref
uv = scenes.msvc_sample_pattern(1)
for sc in (scenes.cornell_box(2048), scenes.load_scene_npz('tests/golden/colorballs.npz')):
    N = sc.numtriangles
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=uv)
    ours = p.runCalculateRadiosityMatrix(0, N, 0)["m_value"]
    for nofma in (True, False):
        ref, secs = pyref.cuda_run_calculate_radiosity_matrix(sc.vertices, sc.normals, sc.tri, nofma=nofma)
        for thr in (1e-12, 1e-9, 1e-7):
            big = np.maximum(ours, ref) > thr
            rel = np.abs(ours - ref)[big] / np.maximum(ours, ref)[big]
            print(sc.name, 'nofma', nofma, 'thr', thr, 'relmax', rel.max(), 'n', big.sum())
        print('   identical frac', (ours == ref).mean(), 'nonzero identical', (ours[ours>0] == ref[ours>0]).mean(), 'facing mismatch', ((ours > 0) != (ref > 0)).sum(), 'ref secs', secs)
        d = np.abs(ours-ref); i = np.unravel_index(np.argmax(np.where(np.maximum(ours,ref)>1e-9, d/np.maximum(np.maximum(ours,ref),1e-300), 0)), d.shape)
        print('   worst', i, ours[i], ref[i])
    p.close()
