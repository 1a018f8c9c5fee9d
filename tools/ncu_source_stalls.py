import csv, subprocess, sys, io
rep, kn = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kn], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(src)))
h = rr[1]
iS, iN, iE, iLS = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed'), h.index('stall_long_sb')
rows = []
tot = 0
for idx, r in enumerate(rr[2:]):
    if len(r) <= iLS or r[0] in ('Kernel Name', 'Address'): break
    try: n = int(r[iN]); ls = int(r[iLS]); ex = int(r[iE])
    except: continue
    rows.append((idx, n, ls, ex, r[iS].strip()))
    tot += n
print('total samples', tot, 'instructions', len(rows))
# print the instruction stream with samples for the top stall sites plus 2 lines of context before
hot = sorted(rows, key=lambda x: -x[1])[:top]
hotidx = sorted(set(i for (i, *_ ) in hot))
for i in hotidx:
    idx, n, ls, ex, s = rows[i]
    print(f'{idx:5d} samp {n:6d} ({100*n/tot:4.1f}%) long_sb {ls:6d} exec {ex:9d}  {s[:90]}')
