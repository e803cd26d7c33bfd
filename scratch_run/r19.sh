python - <<'PY' > gpurun_out/s19_view.log 2>&1
import parmgmc_b200 as pmg
ctx = pmg.Context(0, seed=1)
for dim, dims, lv in ((3, (65, 65, 65), 4), (2, (257, 257, 1), 4)):
    m = pmg.Mat.laplace(ctx, dim, *dims, kappa=1.0)
    pc = pmg.PC(ctx, "gamgmc"); pc.set_operator(m); pc.set_options({"-gamgmc_pc_mg_levels": lv, "-pc_b200_noise": "philox"}); pc.setup()
    print(pc.view())
PY
cat gpurun_out/s19_view.log
