"""ncu --csv (metrics mode, one row per kernel x metric) -> one line per launch with the chosen columns."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
iid, iname, imet, iunit, ival = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
byid = collections.OrderedDict()
for r in rows[1:]:
    d = byid.setdefault(r[iid], {"name": r[iname]})
    v = r[ival].replace(",", "")
    try: v = float(v)
    except ValueError: pass
    d[r[imet]] = (v, r[iunit])
def g(d, k, scale=1.0):
    v = d.get(k)
    if v is None: return float("nan")
    val, unit = v
    if unit in ("Gbyte",): val *= 1e9
    if unit in ("Mbyte",): val *= 1e6
    if unit in ("Kbyte",): val *= 1e3
    if unit in ("us",): val *= 1e-3
    if unit in ("ns",): val *= 1e-6
    return val * scale
print(f"{'id':>4} {'ms':>8} {'dramRdGB':>9} {'rdSect(M)':>10} {'hit%':>6} | {'EL sect(M)':>10} {'EL hit%':>7} | {'EF sect(M)':>10} {'EF hit%':>7} | {'EN sect(M)':>10} {'EN hit%':>7} | {'fabric(M)':>9} {'fab hit%':>8} | {'promoted(M)':>11} {'user miss(M)':>12} {'fill(M)':>8} {'L1hit%':>6}  kernel")
for k, d in byid.items():
    def pair(prefix):
        h, m = g(d, prefix + "_lookup_hit.sum"), g(d, prefix + "_lookup_miss.sum")
        tot = h + m
        return tot / 1e6, (100 * h / tot if tot else float("nan"))
    rd = g(d, "lts__t_sectors_op_read.sum") / 1e6
    rh, rm = g(d, "lts__t_sectors_op_read_lookup_hit.sum"), g(d, "lts__t_sectors_op_read_lookup_miss.sum")
    el, elh = pair("lts__t_sectors_op_read_evict_last")
    ef, efh = pair("lts__t_sectors_op_read_evict_first")
    en, enh = pair("lts__t_sectors_op_read_evict_normal")
    fb, fbh = pair("lts__t_sectors_srcunit_ltcfabric")
    name = d["name"][:70]
    print(f"{k:>4} {g(d,'gpu__time_duration.sum'):8.3f} {g(d,'dram__bytes_read.sum')/1e9:9.2f} {rd:10.1f} {100*rh/(rh+rm) if rh+rm else float('nan'):6.1f} | {el:10.1f} {elh:7.1f} | {ef:10.1f} {efh:7.1f} | {en:10.1f} {enh:7.1f} | {fb:9.1f} {fbh:8.1f} | "
          f"{g(d,'lts__t_sectors_lookup_miss_data_promoted.sum')/1e6:11.1f} {g(d,'lts__t_sectors_lookup_miss_data_user.sum')/1e6:12.1f} {g(d,'lts__d_sectors_fill_device.sum')/1e6:8.1f} {g(d,'l1tex__t_sector_hit_rate.pct'):6.1f}  {name}")
