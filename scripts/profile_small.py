"""One eager training step of a small config (mnist01 | cascade | latent_translator) at its BASELINE batch between
cudaProfilerStart/Stop for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches_cascade.csv python scripts/profile_small.py cascade"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
name = sys.argv[1] if len(sys.argv) > 1 else "cascade"
g = torch.Generator().manual_seed(19)
torch.manual_seed(0)
if name == "mnist01":
    from causal_vae_b200.mnist import models, train
    models.CONFIG["M_DIM"], models.CONFIG["T_DIM"], models.CONFIG["Z_DIM"] = 4, 10, 10
    B = 64
    tr = train.AdversarialTrainer(models.CausalMorphVAE12().cuda(), models.LatentDiscriminator().cuda(), lr=1e-3)
    args = [torch.rand(B, 1, 28, 28, generator=g), torch.rand(B, 4, generator=g), torch.eye(10)[torch.randint(0, 10, (B,), generator=g)]]
elif name == "cascade":
    from causal_vae_b200.cascade import models, train
    B = 256
    tr = train.CascadeTrainer(models.CausalBioVAE(img_channels=1, m_dim=8, t_dim=19, latent_dim=64).cuda(), lr=1e-3)
    args = [torch.randn(B, 1, 64, 64, generator=g), torch.rand(B, 8, generator=g), torch.randint(0, 19, (B,), generator=g),
            torch.randn(B, 64, generator=g)]
else:
    from causal_vae_b200.latent_translator import engine, models
    B = 128
    tr = engine.ViTVAETrainer(models.ViTVAE(img_size=(128, 128)).cuda(), lr=1e-4)
    args = [torch.rand(B, 1, 128, 128, generator=g), torch.randn(B, 512, generator=g)]
args = [a.cuda() for a in args]
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(*args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
