"""DevicePrefetcher: batches arrive in order and intact while copies overlap the consumer's kernels."""
import pytest
import torch



def test_prefetcher_rejects_cpu_device():
    from vit_prisma.b200.prefetch import DevicePrefetcher
    with pytest.raises(RuntimeError):
        DevicePrefetcher([torch.zeros(2)], torch.device("cpu"))


@pytest.mark.gpu
def test_prefetcher_order_and_content():
    from vit_prisma.b200.prefetch import DevicePrefetcher
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    host = [torch.randn(64, 3, 32, 32, generator=g).pin_memory() for _ in range(7)]
    sums = []
    big = torch.randn(4096, 4096, device=dev)
    for x in DevicePrefetcher(host, dev):
        y = big @ big                      # keep the compute stream busy so the next copy really runs underneath
        sums.append((x.double().sum() + 0 * y[0, 0].double()).item())
    assert len(sums) == len(host)
    for s, h in zip(sums, host):
        assert abs(s - h.double().sum().item()) < 1e-6
    # a second pass over fewer batches than the depth
    out = [x.clone() for x in DevicePrefetcher(host[:1], dev)]
    assert len(out) == 1 and torch.equal(out[0].cpu(), host[0])


@pytest.mark.gpu
def test_prefetcher_feed_reuses_buffers_across_passes():
    from vit_prisma.b200.prefetch import DevicePrefetcher
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    host = [torch.randn(8, 16, generator=g).pin_memory() for _ in range(5)]
    pf = DevicePrefetcher(None, dev)
    assert list(pf) == []
    for _ in range(3):
        got = [x.clone() for x in pf.feed(host)]
        assert len(got) == 5 and all(torch.equal(a.cpu(), b) for a, b in zip(got, host))
    ptrs = {b.data_ptr() for b in pf._bufs}
    list(pf.feed(host[:2]))
    assert {b.data_ptr() for b in pf._bufs} == ptrs
