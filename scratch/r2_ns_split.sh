for sp in 0.0 0.3 0.4 0.5 0.6 0.75; do QTB200_OVERLAP_SPLIT=$sp timeout 150 python scratch/r2_ns.py 2>&1 | python -c "
import sys,json
t=sys.stdin.read(); d=json.loads(t[t.index('{'):]); print('$sp', d['code_only_plain']['us'], d['overlapped_equals_plain'], d['overlapped'])"; done
