import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.synthetic import make_config
from oracle import gp_oracle, ocba_oracle
name = "c3"
exp = json.load(open("tests/golden/quoted_configs_expected.json"))[name]
desc, ctx = make_config(name)
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=torch.float64)
arm, x = crs.gao_modellist(model, ctx)
ws = model.workspace(ctx.shape[0])
rows = torch.tensor(exp["rows"])
fix = torch.tensor([float.fromhex(h) for h in exp["row_min_zeta_hex"]], dtype=torch.float64)
got = ws["min_zeta"][rows.cuda()].cpu()
zg, _ = ocba_oracle.zeta_table(ws["mean"].cpu(), ws["var"].cpu())
mu, var = gp_oracle.posterior_grid_batched(desc, ctx, chunk=1024)
zo, _ = ocba_oracle.zeta_table(mu, var)
zg, zo = zg[rows].min(1).values, zo[rows].min(1).values
for i in range(12):
    print(f"row {int(rows[i])}: kernel {float(got[i]):.17e} torch-on-gpu-tables {float(zg[i]):.17e} live-oracle {float(zo[i]):.17e} fixture {float(fix[i]):.17e}")
print("max rel kernel vs gpu-tables", float(((got - zg).abs() / zg).max()), " live oracle vs fixture", float(((zo - fix).abs() / fix).max()),
      " gpu-tables vs live oracle", float(((zg - zo).abs() / zo).max()))
