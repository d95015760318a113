"""ncu --csv metrics table for tools/l2_metrics2.txt: imbalance across L2 slices / DRAM channels, request-level hit rates."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
iid, iname, imet, iunit, ival = [hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")]
byid = collections.OrderedDict()
for r in rows[1:]:
    d = byid.setdefault(r[iid], {"name": r[iname]})
    v = r[ival].replace(",", "")
    try: v = float(v)
    except ValueError: pass
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "us": 1e-3, "ns": 1e-6, "ms": 1.0}.get(r[iunit], 1.0)
    d[r[imet]] = v * mult if isinstance(v, float) else v
g = lambda d, k: d.get(k, float("nan"))
print(f"{'id':>3} {'ms':>7} {'dramGB':>7} {'dram%':>6} {'dram max/avg':>12} {'lts%':>5} {'l1tex%':>6} {'slice max/avg':>13} {'slice min/avg':>13} {'texReq(M)':>9} {'reqHit%':>7} {'texSect(M)':>10} {'sectHit%':>8} {'fabric(M)':>9} {'fabHit%':>7} {'promo(M)':>8} {'user(M)':>8} {'fill(M)':>8} {'L1hit%':>6} {'xbar2l1(M)':>10} {'long_sb':>7}")
for k, d in byid.items():
    rh, rm = g(d, "lts__t_requests_srcunit_tex_lookup_hit.sum"), g(d, "lts__t_requests_srcunit_tex_lookup_miss.sum")
    sh, sm = g(d, "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"), g(d, "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum")
    fh, fm = g(d, "lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum"), g(d, "lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum")
    print(f"{k:>3} {g(d,'gpu__time_duration.sum'):7.3f} {g(d,'dram__bytes_read.sum')/1e9:7.2f} {g(d,'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(d,'dram__bytes_read.max')/g(d,'dram__bytes_read.avg'):12.2f} {g(d,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {g(d,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{g(d,'lts__t_sectors.max')/g(d,'lts__t_sectors.avg'):13.2f} {g(d,'lts__t_sectors.min')/g(d,'lts__t_sectors.avg'):13.2f} "
          f"{g(d,'lts__t_requests_srcunit_tex.sum')/1e6:9.1f} {100*rh/(rh+rm):7.1f} {g(d,'lts__t_sectors_srcunit_tex_op_read.sum')/1e6:10.1f} {100*sh/(sh+sm):8.1f} "
          f"{g(d,'lts__t_sectors_srcunit_ltcfabric.sum')/1e6:9.1f} {100*fh/(fh+fm):7.1f} {g(d,'lts__t_sectors_lookup_miss_data_promoted.sum')/1e6:8.1f} {g(d,'lts__t_sectors_lookup_miss_data_user.sum')/1e6:8.1f} "
          f"{g(d,'lts__d_sectors_fill_device.sum')/1e6:8.1f} {g(d,'l1tex__t_sector_hit_rate.pct'):6.1f} {g(d,'l1tex__m_xbar2l1tex_read_sectors.sum')/1e6:10.1f} {g(d,'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):7.1f}")
