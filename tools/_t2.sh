timeout 600 python -m pytest tests/test_rnn_tc_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -3
python tools/rnn_time.py tensor 2000 32 640 320 2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B32 lstm fwd us/step', round(d['us_per_step_fwd_rec'],3), 'bwd', round(d['us_per_step_bwd_rec'],3))"
B200RNN_TC_PROFILE=1 python tools/rnn_time.py tensor 2000 64 640 320 2 2>&1 | grep "b200rnn" | tail -5 | cut -c1-330
python tools/step_time.py 2>&1 | tail -8
