python tools/sp_accuracy_report.py > gpurun_out/r2l_sp_accuracy.json 2> gpurun_out/r2l_sp.err; tail -3 gpurun_out/r2l_sp.err; cat gpurun_out/r2l_sp_accuracy.json
