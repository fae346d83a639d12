"""Development aid: per-window CTA times of pass 2 (option window_times) for a build + maxWithin."""
import os, re, statistics, subprocess, sys
sys.path.insert(0, '.')
if len(sys.argv) > 3 and sys.argv[3] == "child":
    import pbwt_b200 as P
    M, N = int(sys.argv[1]), int(sys.argv[2])
    eng = P.Engine(0)
    p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
    r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); r.free()
    eng.set_option("window_times", 1)
    r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); r.free()
else:
    M, N = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("10000", "100000")
    out = subprocess.run([sys.executable, __file__, M, N, "child"], capture_output=True, text=True).stderr
    for line in out.splitlines():
        if "pass-2 window times" in line:
            rows = [(int(a), int(b), int(c)) for a, b, c in re.findall(r" (\d+):(\d+)/(\d+)", line)]
            t = [r[1] for r in rows]
            print(len(rows), "windows; first", rows[0], "second", rows[1], "last", rows[-1], "mean", round(statistics.mean(t)), "max", max(t),
                  "top", sorted(rows, key=lambda r: -r[1])[:4])
