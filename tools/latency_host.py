#!/usr/bin/env python3
"""Host-side view of single-proof latency: medians of zkb_fullprove (inputs.json -> proof.json), of the resident
device pass alone, and of the copies around it."""
import json,os,sys,time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from zk_franchise_proof_circuit_b200 import prover
ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
c=prover.load(open(ART+"/proving_key.zkey","rb").read(),open(ART+"/circuit.wasm","rb").read())
doc=open(os.path.join(ROOT, "tests", "golden", "inputs_example.json"),"rb").read()
inp=json.loads(doc)
for _ in range(5): c.fullprove(doc)
def med(f,n=30):
    t=[]
    for _ in range(n):
        t0=time.perf_counter(); f(); t.append((time.perf_counter()-t0)*1e3)
    t.sort(); return round(t[n//2],3)
print("fullprove", med(lambda: c.fullprove(doc)))
print("fullprove_batch[1]", med(lambda: c.fullprove_batch([doc])))
p=np.stack([prover.pack_inputs(inp)])
c.set_inputs(p)
print("set_inputs", med(lambda: c.set_inputs(p)))
print("prove_resident(1)", med(lambda: c.prove_resident(1)))
print("get_results", med(lambda: c.get_results(1)))
print("witness only", med(lambda: c.witness(doc)))
