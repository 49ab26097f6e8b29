"""Randomised comparison of the parallel gzip reader with the serial one (CPU only): a two-member
file is damaged in one of four ways (bit flip, truncation, random bytes, deleted bytes) and read
through both; text or error message must be the same.  usage: fuzz_parallel_gunzip.py [seed] [cases]"""
import gzip, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
TMP = os.path.join(tempfile.mkdtemp(prefix='tb_fuzz_'), 'fz.fq.gz')
from trio_binning_b200 import kmers
import test_parallel_gunzip as T
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
text=T._fastq(60,6000,99)
blob=gzip.compress(text,6)+gzip.compress(text[:50000],1)
os.environ['TB_PGZIP_MIN']='0'
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 200):
    d=bytearray(blob)
    kind=int(rng.integers(0,4))
    if kind==0:
        p=int(rng.integers(0,len(d))); d[p]^=1<<int(rng.integers(0,8))
    elif kind==1:
        p=int(rng.integers(0,len(d))); d=d[:p]
    elif kind==2:
        p=int(rng.integers(0,len(d))); n=int(rng.integers(1,50)); d[p:p+n]=os.urandom(n)
    else:
        p=int(rng.integers(0,len(d))); n=int(rng.integers(1,2000)); del d[p:p+n]
    open(TMP,'wb').write(bytes(d))
    res=[]
    for en,span in (('0',4096),('1',int(rng.integers(1024,60000)))):
        os.environ['TB_PGZIP']=en; os.environ['TB_PGZIP_SPAN']=str(span)
        try: res.append(kmers.fastx_cat(TMP,4))
        except RuntimeError as e: res.append('ERR '+str(e).split(': ',1)[-1])
    if res[0]!=res[1]:
        bad+=1; print('MISMATCH',it,kind,p, res[0][:80] if isinstance(res[0],str) else len(res[0]), res[1][:80] if isinstance(res[1],str) else len(res[1]))
print('done, mismatches',bad)
sys.exit(1 if bad else 0)
