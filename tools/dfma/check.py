import subprocess, sys
Q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
for name, p in (("Fq", Q), ("Fr", R)):
    limbs = [(p >> (52 * i)) & ((1 << 52) - 1) for i in range(5)]
    pinv = (-pow(p, -1, 1 << 52)) % (1 << 52)
    out = subprocess.run(["./proto"] + [hex(x)[2:] for x in limbs] + [hex(pinv)[2:], "200000"], capture_output=True, text=True).stdout
    bad = 0
    for line in out.splitlines():
        v = [int(x, 16) for x in line.split()]
        a = sum(v[i] << (52 * i) for i in range(5)); b = sum(v[5 + i] << (52 * i) for i in range(5)); r = sum(v[10 + i] << (52 * i) for i in range(5))
        if not (r < p and (r << 260) % p == a * b % p):
            bad += 1
    print(name, "checked", len(out.splitlines()), "bad", bad)
