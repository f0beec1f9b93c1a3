"""Static SASS mnemonic counts per kernel of the built library (-> profiles/r2_sass_counts.txt): proves which kernels are
tcgen05 / TMEM / TMA code.  Needs no GPU:  python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-nmf_b200", "lib", "libnmf_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDS", "STS",
        "MUFU.RCP", "MUFU.LG2", "MUFU.EX2", "MUFU.RSQ", "FFMA2", "FFMA", "HMMA")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass pytorch-nmf_b200/lib/libnmf_b200.so : tcgen05 / TMEM / TMA / bulk-copy mnemonics per kernel (static instruction counts)")
print("# UTCHMMA = tcgen05.mma (f16 kind), LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 = packed fp32 pairs; kernels without any tensor / TMA mnemonic are listed with their FFMA count")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for ln in blk.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
                break
    short = name.replace("(int)", "").replace("(bool)", "").replace("nmfb200::<unnamed>::", "").replace("nmfb200::", "")
    print(short.split("(CUtensorMap")[0][:200] if "CUtensorMap" in short else short[:160])
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())) + f"  (instructions: {sum(1 for ln in blk.split(chr(10)) if re.search(r'/[*][0-9a-f]{4}[*]/', ln))})")
