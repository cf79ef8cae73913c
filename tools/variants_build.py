"""Time the single-GPU build and the simulated 8-rank merge for every library variant (development)."""
import os, subprocess, sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
for lib in sorted((ROOT / "regridding_b200" / "variants").glob("lib_*.so")):
    env = dict(os.environ, REGRID_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "prof_sharded.py"), "8"], env=env, capture_output=True, text=True)
    lines = r.stdout.strip().splitlines()
    print(lib.stem, "|", lines[-2][:60] if len(lines) > 1 else r.stderr[-300:], "|", lines[-1][lines[-1].find("merge per band"):][:120] if lines else "")
