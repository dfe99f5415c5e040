"""clock64 timeline of the split-K pair streaming kernel (mlp_i8_umma_pair_kernel), config C5, CTA 0, per layer:
[barrier passed, first operands landed, issuer 0's last MMA issued, accumulators complete, peer's half sent, peer's partial sums
received, outputs stored, published].  NETCUDA_DEBUG_TIMELINE build (NETCUDA_DEBUG_TIMELINE=1 NETCUDA_BUILD_TAG=dbg python build.py;
NETCUDA_LIB_DIR=.../lib_dbg)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["NETCUDA_STREAM_DEBUG_PTR"] = hex(dbg.data_ptr())
import netcuda as nc
np.set_printoptions(linewidth=220)
rng = np.random.default_rng(0)
npl, n_ins = [4096] * 8, 4096
wq = rng.integers(-8, 9, 8 * 4096 * 4096, dtype=np.int8); bq = rng.integers(-2000, 2000, 8 * 4096, dtype=np.int32)
net = nc.Net.mlp(npl, n_ins, precision=nc.PREC_INT8, max_batch=128); net.upload_mlp_i8(wq, bq)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for batch in (17, 128):
    x = torch.randint(-128, 128, (batch, n_ins), dtype=torch.int8, device="cuda"); y = torch.empty((batch, 4096), dtype=torch.int32, device="cuda")
    for _ in range(5): net.forward_device_i8(x, y, batch, s)
    s.synchronize()
    d = dbg.cpu().numpy()[:128].reshape(16, 8)[:8]
    fine = dbg.cpu().numpy()[128:256].reshape(4, 16, 2)
    print("batch", batch, "cycles relative to layer 0's first operands; columns: barrier passed, first operands, issuer 0 done, accumulators complete,"
          " peer half sent, peer sums received, stored, published")
    print(d - d[0, 1])
    dd = d.astype(np.float64)
    print("per layer (cycles): barrier->operands", np.round((dd[1:, 1] - dd[1:, 0]).mean()), " operands->issued", np.round((dd[:, 2] - dd[:, 1]).mean()),
          " issued->complete", np.round((dd[:, 3] - dd[:, 2]).mean()), " complete->sent", np.round((dd[:, 4] - dd[:, 3]).mean()),
          " sent->received", np.round((dd[:, 5] - dd[:, 4]).mean()), " received->stored", np.round((dd[:7, 6] - dd[:7, 5]).mean()),
          " stored->published", np.round((dd[:7, 7] - dd[:7, 6]).mean()), " published->next barrier passed", np.round((dd[1:, 0] - dd[:7, 7]).mean()),
          " layer period", np.round((dd[1:, 1] - dd[:7, 1]).mean()))
    print("layer 3, per issuer and k-block (relative to the layer's barrier-passed stamp): weights landed / activations landed")
    for i in range(4):
        print(" issuer", i, [(int(w - d[3, 0]) if w else None, int(a - d[3, 0]) if a else None) for w, a in fine[i]])
    dbg.zero_()
