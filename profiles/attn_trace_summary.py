"""Turns the stderr of `VP_ATTN_TRACE=1 python profiles/attn_trace.py` into per-step deltas of CTA 0's pipeline."""
import sys
lines = open(sys.argv[1]).read().splitlines()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
blocks = []
for l in lines:
    if l.startswith('attn trace'):
        blocks.append([l])
    elif blocks:
        blocks[-1].append(l)
seen = {}
for b in blocks:
    seen[b[0]] = b          # keep the last launch of each shape
for b in seen.values():
    print(b[0])
    d = {}
    for l in b[1:]:
        p = l.split()
        if len(p) > 5:
            d[p[0]] = [int(x) for x in p[1:]]
    print("step            " + " ".join(f"{i:6d}" for i in range(n)))
    for k, v in d.items():
        if k != '-' and not k.startswith('mma'):
            print(f"{k:16s}" + " ".join(f"{x:6d}" for x in v[:n]))
    for T, u in (("A", 0), ("B", 3)):
        sc, hf, dn, pa = d[f'sm{T}.scores'], d[f'sm{T}.half'], d[f'sm{T}.done'], d[f'sm{T}.p_arrive']
        mp = d['mma.p_full'][u::4]; pv = d['mma.pv_issued'][u::4]; si = d['mma.s_issued'][u::4]
        m = n - 1
        print(f"group {T}{u & 1}: step period (scores[g+1] - scores[g]):", [sc[i + 1] - sc[i] for i in range(m)])
        print(f"group {T}{u & 1}: softmax math (done - scores):", [dn[i] - sc[i] for i in range(m)])
        print(f"group {T}{u & 1}: chain: p_arrive -> issuer sees p_full:", [mp[i] - pa[i] for i in range(m)])
        print(f"group {T}{u & 1}: chain: p_full -> PV issued:", [pv[i] - mp[i] for i in range(m)])
        print(f"group {T}{u & 1}: chain: PV issued -> next S issued:", [si[i + 1] - pv[i] for i in range(m)])
        print(f"group {T}{u & 1}: chain: S issued -> scores seen:", [sc[i] - si[i] for i in range(m)])
        print(f"group {T}{u & 1}: chain total (p_arrive[g] -> scores[g+1]):", [sc[i + 1] - pa[i] for i in range(m)])
