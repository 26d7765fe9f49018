#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_neural_automata.py -x -q > $out/pop_tests.log 2>&1; echo "tests rc=$?"
tail -15 $out/pop_tests.log
timeout 600 python examples/learning_agents.py --epochs 30 --epoch-iters 50 > $out/pop_example.log 2>&1; echo "example rc=$?"
tail -12 $out/pop_example.log
