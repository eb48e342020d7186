"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_step.py into
a per-kernel table of the LAST training step (times are cold-cache and serialised under ncu:
compare shares, not absolutes).   python tools/summarize_launches.py launches.csv [launches/step]"""
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    ours = [r for r in rows if 'tn::' in r['Kernel Name']]
    per = int(sys.argv[2]) if len(sys.argv) > 2 else None
    if per is None:
        # a step starts at the first elastic/warp kernel; find the period from the first name
        first = ours[0]['Kernel Name']
        starts = [i for i, r in enumerate(ours) if r['Kernel Name'] == first]
        per = starts[1] - starts[0] if len(starts) > 1 else len(ours)
    last = ours[-per:]
    unit = last[0]['Metric Unit']
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3}.get(unit, 1e-3)
    tot = sum(float(r['Metric Value'].replace(',', '')) for r in last) * scale
    print('| # | kernel | grid | block | us | share |')
    print('|---|---|---|---|---|---|')
    for i, r in enumerate(last):
        t = float(r['Metric Value'].replace(',', '')) * scale
        name = r['Kernel Name'].split('(')[0].replace('void ', '')
        print('| {} | `{}` | {} | {} | {:.2f} | {:.1%} |'.format(
            i, name, r['Grid Size'].replace(' ', ''), r['Block Size'].replace(' ', ''), t, t / tot))
    print('\n{} launches per step, {:.1f} us summed (serialised, cold cache)'.format(per, tot))


if __name__ == '__main__':
    main()
