#!/bin/bash
# compute-sanitizer over the parity tests of the kernels with shared-memory reductions, tickets and staging rings
# usage (under gpurun): bash tools/sanitize.sh   (round 1, r01t: memcheck / racecheck / synccheck all 0 errors)
# (the peer-memory tail is left out: its ranks wait for each other and the sanitizer serialises kernels)
K1="decode_select or ragged or grad_check or style_loss_golden or ring_depths or alpha_feed or student_step_golden or mean_std_backward or loader_side or both_directions or wide_route or sgd_resumes or pair_route or gather_decode_equals"
K2="decode_select_equals_decode_then_select and C1 or style_loss_golden or ring_depths or grad_check or wide_route or gather_decode_equals and 32-16"
K3="decode_select_equals_decode_then_select and C5 or style_loss_golden or ring_depths or gather_decode_equals and 3-21"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 5 python -m pytest tests -m gpu -x -q -k "$K1" 2>&1 | tail -4
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 5 python -m pytest tests -m gpu -x -q -k "$K2" 2>&1 | tail -4
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 99 --print-limit 5 python -m pytest tests -m gpu -x -q -k "$K3" 2>&1 | tail -4
