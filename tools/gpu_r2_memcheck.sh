#!/bin/bash
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=60000
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x --timeout 1400 \
  "tests/test_gpu_level.py::test_level_orders_bit_exact" "tests/test_gpu_level.py::test_level_staged_is_chosen" \
  "tests/test_gpu_level.py::test_level_replicas_per_word_bit_exact" \
  "tests/test_gpu_colour.py::test_resident_integer_kernel_bit_exact" "tests/test_gpu_colour.py::test_resident_kernel_periodic_trotter_and_orders" \
  "tests/test_gpu_colour.py::test_qa_carry_bit_exact" "tests/test_gpu_colour.py::test_energy_histogram_on_device" > gpurun_out/memcheck_r2.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck_r2.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x --timeout 800 \
  "tests/test_gpu_level.py::test_level_staged_is_chosen" "tests/test_gpu_colour.py::test_resident_kernel_periodic_trotter_and_orders" > gpurun_out/racecheck_r2.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck_r2.log | tail -6
