set -x
N=${1:-2}
python -m pytest tests/test_job.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest_job_n$N.log
python tools/job_check.py $N > gpurun_out/r2c_job_check_n$N.txt 2> gpurun_out/r2c_job_check_n$N.err
tail -5 gpurun_out/r2c_job_check_n$N.err
