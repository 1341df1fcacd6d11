import sys, time, os, tempfile, random
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import owc_b200
from owc_b200 import eval_metrics as EM
import pandas as pd
from concurrent.futures import ProcessPoolExecutor
import multiprocessing as mp
if __name__ == "__main__":
    d = tempfile.mkdtemp()
    files=[]
    for m in range(32):
        rows=[{"doc_id": i, "doc": {"image": f"img_{i}.jpg", "target": "class name %d" % (i%100)}, "target": "class name %d" % (i%100),
               "arguments": [["What type of object is in this photo?", {"max_new_tokens": 32, "temperature": 0}]],
               "resps": [["a photo of thing %d" % i]], "filtered_resps": ["thing %d word" % i], "doc_hash": "x"*64, "prompt_hash": "y"*64, "target_hash": "z"*64,
               "input": "What type of object is in this photo?"} for i in range(8000)]
        f=os.path.join(d, f"m{m}_samples_t.jsonl"); pd.DataFrame(rows).to_json(f, lines=True, orient="records"); files.append(f)
    print("file size MB", os.path.getsize(files[0])/1e6)
    cols={"semantic_similarity":[0.5]*8000}
    t=time.perf_counter()
    for f in files:
        EM._worker_load(f); EM._worker_write((f, cols))
    ts=time.perf_counter()-t
    t=time.perf_counter()
    with ProcessPoolExecutor(max_workers=16, mp_context=mp.get_context("spawn")) as pool:
        list(pool.map(EM._worker_load, files)); list(pool.map(EM._worker_write, [(f, cols) for f in files]))
    tp=time.perf_counter()-t
    print(f"pandas: serial {ts:.2f}s  pool(16, incl. spawn) {tp:.2f}s for {len(files)} files of 8000 rows")
    # native scanner (csrc/jsonl.cpp): serial, then 16 threads (the scanner runs outside the GIL)
    from concurrent.futures import ThreadPoolExecutor
    def native(f):
        raw, items = EM._load_items_native(f)
        EM._write_columns_native(f, raw, cols)
    t=time.perf_counter()
    for f in files: native(f)
    tn=time.perf_counter()-t
    t=time.perf_counter()
    with ThreadPoolExecutor(16) as tp_: list(tp_.map(native, files))
    tnt=time.perf_counter()-t
    print(f"native: serial {tn:.2f}s  threads(16) {tnt:.2f}s")
