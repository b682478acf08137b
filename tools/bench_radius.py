"""Times the two radius searches (O(n m) warp sweep vs cell-binned) at the BASELINE cfg-5 size (N = 262144 points, 32^3 latent
tokens, r = 0.07) and at cfg-2; both directions (encoder: queries = latent, decoder: queries = points); checks equality."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, synth  # noqa: E402

for tag, d, n, lat_size, r in (("cfg2", 2, 9216, (64, 64), 0.033), ("cfg5", 3, 262144, (32, 32, 32), 0.07)):
    x = synth.uniform_points(n, d, 7).cuda()
    lat = synth.latent_grid(lat_size).cuda()
    for name, dd, qq in (("enc", x, lat), ("dec", lat, x)):
        res = {}
        for method in ("sweep", "grid"):
            ops.radius_csr(dd, qq, r, method=method)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                g = ops.radius_csr(dd, qq, r, method=method)
            torch.cuda.synchronize()
            res[method] = ((time.perf_counter() - t0) / 3 * 1e3, g)
        same = torch.equal(res["sweep"][1]["neighbors_index"], res["grid"][1]["neighbors_index"]) and \
            torch.equal(res["sweep"][1]["neighbors_row_splits"], res["grid"][1]["neighbors_row_splits"])
        print(f"{tag} {name}: E = {res['grid'][1]['neighbors_index'].numel()}  sweep {res['sweep'][0]:8.2f} ms  grid {res['grid'][0]:8.2f} ms  "
              f"speed-up {res['sweep'][0] / res['grid'][0]:6.1f}x  identical = {same}")
