import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import helpers as H, oracle_lib as O, ref_witness as RW
from zk_franchise_proof_circuit_b200 import prover
ART=H.ART
c=prover.load(open(ART+'/proving_key.zkey','rb').read(), open(ART+'/circuit.wasm','rb').read())
inp=H.fixture_inputs()
c.set_blinding(H.R_FIXED,H.S_FIXED)
pj,sj=c.fullprove(json.dumps(inp))
got_json=O.proof_bin(json.loads(pj))
c.set_inputs(np.stack([prover.pack_inputs(inp)]))
c.prove_resident(1)
pr,pu,st=c.get_results(1)
got_res=pr[0].tobytes()
code,w=RW.witness(inp)
exp,part=H.zkey_ref().prove(w,H.R_FIXED,H.S_FIXED,partials=True)
def show(name,b):
    print(name, ' '.join(b[i:i+32][:6].hex() for i in range(0,256,32)))
show('json',got_json); show('res ',got_res); show('exp ',exp)
print('pj',pj[:120])
