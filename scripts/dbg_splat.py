import sys, torch
sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat
torch.manual_seed(0)
for (c,h,w,mode) in [(1,64,96,"sum"),(64,64,96,"soft"),(64,1152,1920,"soft")]:
    x=torch.randn((1,c,h,w),device="cuda"); flow=torch.full((1,2,h,w),3.3,device="cuda"); m=torch.randn((1,1,h,w),device="cuda") if mode=="soft" else None
    y=softsplat(x,flow,m,mode,_variant=4); torch.cuda.synchronize(); print("ok",c,h,w,mode,float(y.abs().sum()))
