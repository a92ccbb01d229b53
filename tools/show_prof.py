import json, sys
d = json.load(open(sys.argv[1]))
n = d.get('steps', 2)
tot = sum(r['ms'] for r in d['rows']) / n
print("total kernel ms/step (top rows): %.2f" % tot)
for r in d['rows'][:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-24s %-40s n=%3d ms/step=%7.3f GB/s=%7.0f TF=%6.2f" % (
        r['kernel'][6:], r['shape'], r['launches'] // n, r['ms'] / n,
        r['bytes'] / r['ms'] / 1e6 if r['ms'] else 0, r['flops'] / r['ms'] / 1e9 if r['ms'] else 0))
