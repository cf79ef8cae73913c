"""Build tuning variants of the library into regridding_b200/variants/ (development).
usage: python tools/build_variants.py name=DEF1,DEF2 ..."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from regridding_b200 import _build
out = ROOT / "regridding_b200" / "variants"
out.mkdir(exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    defs = [d for d in defs.split(",") if d]
    _build.build(defines=defs, out=out / f"lib_{name}.so")
    print("built", name, defs)
