"""Sum a B200_TRACE_FILE launch trace by tag: device microseconds and launches per tag (development aid)."""
import sys, csv, collections
for path in sys.argv[1:]:
    dev, cnt = collections.Counter(), collections.Counter()
    with open(path) as f:
        for line in f:
            row = line.rstrip("\n").split(",")
            if len(row) < 4 or not row[0].isdigit():
                continue
            tag = ",".join(row[1:-2])          # tags may contain commas (template arguments)
            dev[tag] += float(row[-2]); cnt[tag] += 1
    tot = sum(dev.values())
    print(f"== {path}: {tot / 1e3:.1f} ms device time, {sum(cnt.values())} launches")
    for tag, us in dev.most_common(24):
        print(f"  {tag:28s} {cnt[tag]:6d} x {us / cnt[tag]:9.1f} us = {us / 1e3:9.1f} ms")
