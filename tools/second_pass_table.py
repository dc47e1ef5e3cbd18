"""ms, DRAM bytes and GB/s per second-pass kernel launch from the ncu CSV tools/r02_second_pass*.sh write: second_pass_table.py CSV"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; d = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        x = dict(zip(hdr, r)); key = (x['ID'], x['Kernel Name'].split('(')[0][:64])
        d.setdefault(key, {})[x['Metric Name']] = (float(x['Metric Value'].replace(',', '')), x['Metric Unit'])
sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3}
for (i, k), m in d.items():
    t = m['gpu__time_duration.sum'][0] * sc[m['gpu__time_duration.sum'][1]]
    rd = m['dram__bytes_read.sum'][0] * sc[m['dram__bytes_read.sum'][1]]; wr = m['dram__bytes_write.sum'][0] * sc[m['dram__bytes_write.sum'][1]]
    print("%-66s %8.3f ms  read %6.3f GB write %6.3f GB -> %6.0f GB/s DRAM (%.2f of 6538.6)" % (k, t * 1e3, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e9, (rd + wr) / t / 1e9 / 6538.6))
