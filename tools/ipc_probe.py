import torch, torch.multiprocessing as mp, time
def child(q, q2):
    torch.cuda.set_device(1)
    t = q.get()            # tensor living on cuda:0, opened through cudaIpcOpenMemHandle
    print("child got", t.device, t.shape, flush=True)
    src = torch.arange(t.numel(), dtype=torch.float64, device="cuda:1")
    torch.cuda.synchronize()
    t0=time.time()
    for _ in range(10): t.copy_(src)          # peer write over NVLink by the copy engine
    torch.cuda.synchronize(); dt=(time.time()-t0)/10
    print("peer copy GB/s", t.numel()*8/dt/1e9, flush=True)
    q2.put("done")
if __name__ == "__main__":
    mp.set_start_method("spawn")
    print("peer access 0->1", torch.cuda.can_device_access_peer(0,1))
    q, q2 = mp.Queue(), mp.Queue()
    p = mp.Process(target=child, args=(q,q2)); p.start()
    torch.cuda.set_device(0)
    t = torch.zeros(64*1024*1024, dtype=torch.float64, device="cuda:0")
    q.put(t)
    print(q2.get(timeout=120))
    torch.cuda.synchronize()
    print("check", float(t[12345]), float(t[-1]))
    p.join()
