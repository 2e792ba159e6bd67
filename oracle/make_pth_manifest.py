"""Generate tests/golden/pth_manifest.json FROM THE REFERENCE's released checkpoints (run here, where /root/reference
exists):   python -m oracle.make_pth_manifest

For every pth/*.pth: the ordered list of (key, shape, dtype) and the fraction of exactly-zero weights.  The weights
themselves are committed only for the files the parity tests run (tests/golden/ckpt); the manifest pins the
state_dict LAYOUT of all eighteen, which is what a drop-in for trainer.py / tester.py / detect.py has to accept.
Also tests/golden/netcfg_manifest.json: the three weights*/net.cfg layer lists, parsed (export.net_cfg must emit them)."""
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


def net_cfgs():
    """tests/golden/netcfg_manifest.json: the reference's three hand-written layer lists (weights*/net.cfg), parsed
    into [section, [[key, value], ...]] plus the sha256 of the text without trailing blank lines."""
    import hashlib

    def parse_net_cfg(text):  # own parser: the fixture must not be computed by the code it checks
        out = []
        for raw in text.splitlines():
            ln = raw.strip()
            if not ln or ln.startswith("#"):
                continue
            if ln.startswith("["):
                out.append((ln.strip("[]").strip(), []))
            else:
                k, v = ln.split("=", 1)
                out[-1][1].append((k.strip(), v.strip()))
        return out

    ref = REF.parent
    man = {}
    for d in ("weights", "weightsVGA", "weightsLP"):
        text = (ref / d / "net.cfg").read_text()
        man[d] = {"sections": [[sec, [list(kv) for kv in kvs]] for sec, kvs in parse_net_cfg(text)],
                  "sha256_rstrip": hashlib.sha256(text.rstrip().encode()).hexdigest()}
    out = OUT.parent / "netcfg_manifest.json"
    out.write_text(json.dumps(man, separators=(",", ":")))
    print(f"net.cfg x{len(man)} -> {out} ({out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
    net_cfgs()
