"""Time config 5 (tools/prof_locate.py) for every library variant under regridding_b200/variants (development)."""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
for lib in sorted((ROOT / "regridding_b200" / "variants").glob("lib_*.so")):
    env = dict(os.environ, REGRID_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "prof_locate.py")], env=env, capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if "ms" in l]
    print(lib.stem, "|", " | ".join(l[:70] + l[l.find("equals"):] for l in lines) if lines else r.stderr[-300:])
