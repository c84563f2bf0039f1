"""For EVERY registration of the five in-scope families: builds the reference's own Keras model (unmodified
reference code on the TF shim, oracle/ref_runner.py), and checks that its variables (names + shapes, minus the
non-loadable buffers) are exactly the oracle's ``param_shapes`` and the engine's ``param_specs``.
Writes profiles/r02_registrations_vs_reference.txt.   python tools/check_registrations_vs_reference.py
"""
import importlib
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))

import tfimm  # noqa: E402
from oracle import ref_runner as rr  # noqa: E402

IGNORE = ("attn_mask", "relative_position_index", "blur_kernel")


def main():
    lines, bad = [], 0
    t_all = time.time()
    for fam in rr.FAMILIES:
        omod = importlib.import_module(f"oracle.{fam}")
        for name in rr.list_models(module=fam):
            t0 = time.time()
            ref = rr.create_model(name)
            rs = {k: v for k, v in ref.weight_shapes().items() if not any(p in k for p in IGNORE)}
            cfg = tfimm.models.model_config(name)
            os_ = {k: tuple(v) for k, v in omod.param_shapes(cfg).items()}
            eng = tfimm.create_model(name, device="meta")
            es = {k: tuple(s.shape) for k, s in eng.param_specs().items() if not any(p in k for p in IGNORE)}
            ok = rs == os_ == es
            bad += not ok
            n = sum(int(__import__("numpy").prod(s)) for s in rs.values())
            lines.append(f"{'OK ' if ok else 'BAD'} {fam:13s} {name:45s} {len(rs):4d} variables {n:>12,d} params "
                         f"{time.time() - t0:5.1f}s")
            print(lines[-1], flush=True)
            del ref
    lines.append(f"{len(lines)} registrations, {bad} mismatches, {time.time() - t_all:.0f}s")
    (ROOT / "profiles" / "r02_registrations_vs_reference.txt").write_text("\n".join(lines) + "\n")
    print(lines[-1])
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
