# throughput sweep of the residency knobs on the hall path (us per frame with N frames in flight)
cd $GRAFT_REPO_ROOT
run() { env "$@" timeout 200 python profiles/ab.py hall 384 $FL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['knobs'], 'in_flight', $FL, d['us_per_frame_${FL}_in_flight'])"; }
for FL in 8 12 16 24; do run SRB_DUMMY=1; done
FL=12
for r in 2 3 4 5; do run SRB_RASTER_CTAS_PER_SM=$r; done
for s in 4 6 8 10 12; do run SRB_SHADE_CTAS_PER_SM=$s; done
for s in 1 2 3; do run SRB_SETUP_CTAS_PER_SM=$s; done
run SRB_RASTER_CTAS_PER_SM=4 SRB_SHADE_CTAS_PER_SM=8
run SRB_RASTER_CTAS_PER_SM=2 SRB_SHADE_CTAS_PER_SM=8
FL=16; run SRB_RASTER_CTAS_PER_SM=2 SRB_SHADE_CTAS_PER_SM=4
