"""Research: dump the correlation matrix of the bench's (torch-generated) instance for CPU studies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import bench
from macrodna_b200 import get_handle, _lib

h = get_handle(0)
dev = torch.device("cuda", 0)
for wl in sys.argv[1:]:
    M, N, G, clones = bench.SHAPES[wl]
    rna, dna, rc, dc = bench.make_device_instance(torch, M, N, G, clones, 1234 + int(wl[1:]), dev)
    corr = np.empty((M, N))
    a, s, o, st = h.cell2cell(rna.data_ptr(), dna.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE, corr_out=corr)
    np.savez_compressed("gpurun_out/corr_torch_%s.npz" % wl, corr=corr.astype(np.float64), assign=a, step=s, obj=o,
                        rna_clone=rc.cpu().numpy(), dna_clone=dc.cpu().numpy())
    print(wl, st.as_dict()["step_rounds"], o)
