"""Diagnostic: row-resident forward sweep (gru_rows.cu) vs the slice kernels on the same encoder forward (B = 2048)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vame_oracle as vo
from vame_b200.engine import Engine
from vame_b200 import _lib as L

lib = L.lib()
T, F, Z, H, B = int(os.environ.get("T", 4)), 24, 30, 256, int(os.environ.get("B", 2048))
torch.manual_seed(19)
port = vo.RefPort(2 * T, Z, F, False, 0, hidden=H)
eng = Engine(F, T, Z, H, H, H, False, 0, False, device="cuda")
eng.load_state_dict(port.state_dict())
x = torch.randn(B, T, F, generator=torch.Generator().manual_seed(1))
outs = {}
for rows in (1, 0):
    lib.vame_set_option(b"rows", rows)
    outs[rows] = eng.encoder_forward(x.cuda()).cpu()
with torch.no_grad():
    ref = port.encode(x)
for name, a in (("rows", outs[1]), ("slice", outs[0])):
    for i, piece in enumerate(("L0 fwd", "L0 bwd", "L1 fwd", "L1 bwd")):
        d = (a[:, i * H:(i + 1) * H] - ref[:, i * H:(i + 1) * H]).abs()
        print("%-5s %-6s max abs err %.3e (ref max %.3f)  worst unit %d row %d" % (name, piece, float(d.max()), float(ref[:, i * H:(i + 1) * H].abs().max()),
                                                                               int(d.max(0).values.argmax()), int(d.max(1).values.argmax())))
d = (outs[1][:, :H] - ref[:, :H]).abs()
print("per-32-unit-slice max err (L0 fwd):", [round(float(d[:, 32 * c:32 * c + 32].max()), 5) for c in range(8)])
print("per-128-row-tile max err (L0 fwd):", [round(float(d[128 * t:128 * t + 128].max()), 5) for t in range(min(B // 128, 8))])
