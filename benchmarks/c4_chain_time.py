import sys, json
sys.path.insert(0, "/root/repo")
from benchmarks import config_lines as CL
out = CL.c4_bayes(6540.8)
print(json.dumps({k: v for k, v in out.items() if 'gibbs_200' in k or 'per_s' in k and not isinstance(v, dict)}, indent=1))
