python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 --no-python tests/cpp/_build/test_dist 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | head -40
tests/cpp/_build/test_dist
tests/cpp/_build/test_dropin_reducers
