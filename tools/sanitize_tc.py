"""Small runs of the three tensor-core kernels for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_tc.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import selenite_lite_b200 as slb
for chain, T in ((slb.CHAIN_RX_SSB_F32, 768 * 2 + 384), (slb.CHAIN_TX_SSB_F32, 768 * 2 + 384), (slb.CHAIN_RX_SSB_Q15, 768 * 2 + 48 * 5)):
    C = 11
    x = torch.from_numpy(slb.synth_iq(C, T)).cuda()
    d = slb.DspIf(C, chain=chain)
    y = d.rx_process(x, _dir="tx") if chain == slb.CHAIN_TX_SSB_F32 else d.rx_process(x)
    y2 = d.rx_process(x[:, :768].contiguous(), _dir="tx") if chain == slb.CHAIN_TX_SSB_F32 else d.rx_process(x[:, :768].contiguous())
    torch.cuda.synchronize()
    print("chain", chain, "ok", int(y.abs().sum()), d.kernel_launches())
