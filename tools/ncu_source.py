"""Per-region and per-instruction warp-stall samples of one kernel from an .ncu-rep (run here, no GPU needed):
python tools/ncu_source.py gpurun_out/x.ncu-rep [first last]   -> regions by execution count; with a range, the instructions"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    S = [int(r[ix['# Samples']]) for r in data]
    E = [int(r[ix['Instructions Executed']]) for r in data]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(S)
    print('kernel', rows[0][1][:80], 'samples', tot)
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    print({k.replace('stall_', ''): v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
    if len(sys.argv) >= 4:
        for i in range(int(sys.argv[2]), int(sys.argv[3])):
            r = data[i]
            st = {s.replace('stall_', ''): int(r[ix[s]]) for s in stalls if int(r[ix[s]] or 0) > 0 and s != 'stall_selected'}
            print(i, r[ix['Source']].strip()[:64].ljust(64), str(S[i]).rjust(6), str(E[i]).rjust(9), st)
        return
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and E[j + 1] == E[i]:
            j += 1
        s = sum(S[i:j + 1])
        if s > tot / 300:
            print(f'{i:5d}-{j:5d} exec {E[i]:9d} instrs {j - i + 1:5d} samples {s:7d} {100 * s / tot:5.1f}%')
        i = j + 1


if __name__ == '__main__':
    main()
