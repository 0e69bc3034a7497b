"""Print the measured FP32 FFMA / packed FFMA2 throughput of cuda:0 (roofline denominators)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import engine
torch.cuda.init()
print("FFMA  TFLOP/s:", engine.probe_ffma_tflops(0))
print("FFMA2 TFLOP/s:", engine.probe_ffma_tflops(0, packed=True))
