cd $GRAFT_REPO_ROOT
for sc in cubes rand; do for fl in 4 8 12 16; do timeout 200 python profiles/ab.py $sc 192 $fl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$sc', $fl, d['us_per_frame_${fl}_in_flight'], d['us_per_frame_1_in_flight'])"; done; done
