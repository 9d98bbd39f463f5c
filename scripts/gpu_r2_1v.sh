cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-dqn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), {k:round(v['us_per_ply'],1) for k,v in d['aux']['api_mode'].items()}, 'c5 %.3e' % d['aux']['config5_1M_envs_steps_per_s'])"
