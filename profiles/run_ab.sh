# usage: bash profiles/run_ab.sh <scene> <frames> "<knobs A>" "<knobs B>" ...   ("-" = no knobs)
cd $GRAFT_REPO_ROOT
scene=$1; frames=$2; shift 2
for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then env timeout 300 python profiles/ab.py $scene $frames 2>&1 | tail -2
  else env $cfg timeout 300 python profiles/ab.py $scene $frames 2>&1 | tail -2; fi
done
