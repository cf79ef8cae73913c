"""Time the single-GPU build and its two walks for every library variant under regridding_b200/variants (development)."""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
for lib in sorted((ROOT / "regridding_b200" / "variants").glob("lib_*.so")):
    env = dict(os.environ, REGRID_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "prof_band_trace.py"), "1", "0", "3"], env=env, capture_output=True, text=True)
    keep = [l.strip() for l in r.stdout.splitlines() if "k_walk" in l or "span" in l]
    print(lib.stem, "|", " | ".join(keep) if keep else r.stderr[-300:])
