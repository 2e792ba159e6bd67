"""Generate tests/golden/pth_manifest.json FROM THE REFERENCE's released checkpoints (run here, where /root/reference
exists):   python -m oracle.make_pth_manifest

For every pth/*.pth: the ordered list of (key, shape, dtype) and the fraction of exactly-zero weights.  The weights
themselves are committed only for the files the parity tests run (tests/golden/ckpt); the manifest pins the
state_dict LAYOUT of all eighteen, which is what a drop-in for trainer.py / tester.py / detect.py has to accept."""
import json
from pathlib import Path

import torch

REF = Path("/root/reference/pth")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "pth_manifest.json"


def main():
    man = {}
    for f in sorted(REF.glob("*.pth")):
        sd = torch.load(f, map_location="cpu", weights_only=False)
        w = [v for k, v in sd.items() if v.dim() > 1]
        man[f.stem] = {
            "entries": [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()],
            "params": int(sum(v.numel() for v in sd.values())),
            "zero_weight_fraction": round(float(sum(int((v == 0).sum()) for v in w)) / max(1, sum(v.numel() for v in w)), 4),
        }
    OUT.write_text(json.dumps(man, indent=0, separators=(",", ":")))
    print(f"{len(man)} checkpoints -> {OUT} ({OUT.stat().st_size} bytes)")
    for k, v in man.items():
        print(f"  {k}: {len(v['entries'])} entries, {v['params']} values, {v['zero_weight_fraction']:.3f} zero weights")


if __name__ == "__main__":
    main()
