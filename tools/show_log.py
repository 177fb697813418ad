import json, sys
def fmt(v):
    if isinstance(v, dict): return '{' + ', '.join(f'{k}:{fmt(x)}' for k, x in v.items()) + '}'
    if isinstance(v, float): return f'{v:.4g}'
    return str(v)[:600]
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith('{'):
        print(line[:300]); continue
    d = json.loads(line)
    print(d.pop('check'), ' '.join(f'{k}={fmt(v)}' for k, v in d.items()))
