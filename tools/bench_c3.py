import sys, json
sys.path.insert(0, '.')
import torch, nca_b200
exec(open('tools/bench_configs.py').read().split("for prec in")[0].replace('import sys as _s0',''))
dynca("c3 CD 256x256 C12 fc96 edges circular B64 T80", nca_b200.DyNCA_CD, 64, 12, 96, 256, 256, 80, "bf16", padding_mode="circular", conditioning="edges", edge_transform="None")
dynca("c1 EC 128x128 C12 fc96 CPE B4 T64", nca_b200.DyNCA_EC, 4, 12, 96, 128, 128, 64, "bf16", padding_mode="replicate", pos_emb="CPE")
