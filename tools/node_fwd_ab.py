"""Bitwise A/B of the forward node kernels: prints a digest of encoder / decoder outputs for a few shapes; run it once with
GJ_NODE_FWD_V1=1 (CTA-tile kernels) and once without (warp-autonomous kernels) -- the digests must be identical."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_jet_autoencoder_b200.config import build_models
from gnn_jet_autoencoder_b200.trainer import synthetic_jets

out = []
for prec in ("fp32", "bf16"):
    for (B, N) in ((64, 30), (5, 7), (37, 33), (3, 150)):
        enc, dec = build_models(N, device="cuda", precision=prec, seed=3)
        x = torch.from_numpy(synthetic_jets(B, N, seed=11)).cuda()
        with torch.no_grad():
            z = enc(x)
            y = dec(z)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
        out.append(hashlib.sha1(z.cpu().numpy().tobytes() + y.cpu().numpy().tobytes()).hexdigest()[:16])
print(" ".join(out))
