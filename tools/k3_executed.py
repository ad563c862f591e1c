"""Executed warp-instructions per region of the K3 kernels from an `ncu --set full --import-source on` report
(`ncu --page source --csv`).  Usage: python tools/k3_executed.py REPORT.ncu-rep ID_K3A ID_K3C > profiles/<tag>_k3_executed.txt
(ID_* = the report's kernel ids of the first-scale kernel and of one streaming-kernel launch; run where ncu is installed)."""
import collections, csv, os, subprocess, sys, tempfile

REP, ID_A, ID_C = sys.argv[1], sys.argv[2], sys.argv[3]
W, H, L = 2048, 2048, 512
TXC = 124                     # tile width of the streaming kernel
tmp = tempfile.mkdtemp()


def load(kid):
    path = os.path.join(tmp, f"k{kid}.csv")
    subprocess.run(f"ncu -i {REP} --page source --csv --kernel-id :::{kid} > {path} 2>/dev/null", shell=True)
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hi]
    iex = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); iw = hdr.index('L1 Wavefronts Shared'); ismp = hdr.index('# Samples')
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] == 'Kernel Name':
            break
        try:
            n = int(r[iex])
        except Exception:
            continue
        body.append((n, r[isrc].strip(), int(r[iw] or 0), int(r[ismp] or 0)))
    return rows[0][1], body


def hist(body, lo, hi, div):
    c = collections.Counter()
    for n, s, wv, sm in body[lo:hi + 1]:
        tk = s.split(); op = tk[1] if tk[0].startswith('@') else tk[0]
        c[op.split('.')[0]] += n
    return sum(c.values()) / div, '  '.join(f'{k}:{v / div:.1f}' for k, v in c.most_common(24))


print(f"# Executed warp-instructions per region of the K3 kernels, from `ncu --page source --csv` of {os.path.basename(REP)}")
print(f"# ({W}x{H}x{L}, sigma 2,4,6, fma smoothing).  wp = warp-plane: one warp's iteration for one z plane")
print(f"# ({TXC} voxels in the streaming kernel K3a', 128 x 2 rows in the first-scale kernel K3a).")
name, body = load(ID_C)
tot = sum(b[0] for b in body)
ntx = (W - 4 + TXC - 1) // TXC
WP = ntx * (H // 8) * 8 * L
vox = W * H * L
print(f"\n== {name}\n   total {tot:.4g} warp-inst = {tot / WP:.1f} per wp = {tot * 32 / vox:.1f} thread-inst/voxel; shared-memory wavefronts {sum(b[2] for b in body) / WP:.1f} per wp")
idx_arr = [i for i, b in enumerate(body) if 'SYNCS.ARRIVE' in b[1] and b[0] > WP * 0.9]
a_end = idx_arr[-1]
print("   region, warp-inst per wp, share of samples, opcode mix per wp")


def reg(title, lo, hi):
    t, h = hist(body, lo, hi, WP); s = sum(b[3] for b in body[lo:hi + 1]) / max(1, sum(b[3] for b in body))
    print(f"   {title}: {t:.1f}/wp, {100 * s:.1f}% of samples\n      {h}")


first_lds = next(i for i, b in enumerate(body) if b[1].startswith('LDS.128') and b[0] > WP * 0.5)
reg("loop top (issue duty, full-barrier wait incl. spin, ring offsets)", 0, first_lds - 1)
reg("phase A (second differences, test, append, release)", first_lds, a_end)
reg("phase B (drain: queue read, J gather, eigen, vesselness, update) + out-of-line wait loops", a_end + 1, len(body) - 1)
name, body = load(ID_A)
tot = sum(b[0] for b in body); Q = vox / 128
print(f"\n== {name}\n   total {tot:.4g} warp-inst = {tot / Q:.1f} per warp-quad (128 voxels) = {tot * 32 / vox:.1f} thread-inst/voxel; shared-memory wavefronts {sum(b[2] for b in body) / Q:.1f} per warp-quad")
t, h = hist(body, 0, len(body) - 1, Q)
print("   opcode mix per warp-quad:\n      " + h)
