"""GPU: the hand-written tcgen05 GEMM (csrc/tc_gemm.cu) against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands. Tolerance: fp32 accumulation of bf16 products ->
|err| <= 2e-3 * sqrt(K) * scale for f32 outputs; bf16 outputs add one bf16 rounding (2^-8 relative)
and the tanh.approx error (2^-11)."""
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


# (M >= 18 944 with N a multiple of 256 takes the CTA-pair kernel, tcgen05 cta_group::2: ragged M, one and two
# column tiles, every K of the MLP)
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 64), (4096, 512, 256), (300, 256, 512), (131072, 512, 512),
                                   (1000, 64, 256), (19001, 256, 64), (40000, 512, 256), (131072, 256, 512),
                                   (18944, 512, 512)])
def test_forward_bias_tanh(M, N, K):
    from rsoccer_isaac_cleanrl_b200.engine import EPI_BIAS_TANH_BF16, gemm_bf16
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    gemm_bf16(a, w, out, EPI_BIAS_TANH_BF16, bias=bias)
    ref = torch.tanh(a.float() @ w.float().t() + bias)
    err = (out.float() - ref).abs().max().item()
    assert err < 1.2e-2, err


def test_bias_f32_and_padded_input():
    from rsoccer_isaac_cleanrl_b200.engine import EPI_BIAS_F32, gemm_bf16
    M, N = 1024, 256
    x = torch.zeros((M, 64), device="cuda", dtype=torch.bfloat16)   # 52 observation columns padded to 64
    x[:, :52] = _rand((M, 52), 3)
    w = torch.zeros((N, 64), device="cuda", dtype=torch.bfloat16)
    w[:, :52] = _rand((N, 52), 4, 52 ** -0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.empty((M, N), device="cuda")
    gemm_bf16(x, w, out, EPI_BIAS_F32, bias=bias)
    ref = x[:, :52].float() @ w[:, :52].float().t() + bias
    assert (out - ref).abs().max().item() < 2e-3


def test_dgrad_fused_tanh_derivative():
    from rsoccer_isaac_cleanrl_b200.engine import EPI_DTANH_BF16, gemm_bf16
    M, N_out, K_in = 2048, 512, 256
    dz, wt = _rand((M, N_out), 5), _rand((K_in, N_out), 6, N_out ** -0.5)   # wt = W^T, [K_in, N_out]
    y = torch.tanh(_rand((M, K_in), 7).float()).to(torch.bfloat16)          # activations feeding this layer
    out = torch.empty((M, K_in), device="cuda", dtype=torch.bfloat16)
    gemm_bf16(dz, wt, out, EPI_DTANH_BF16, aux=y)
    ref = (dz.float() @ wt.float().t()) * (1 - y.float() ** 2)
    assert (out.float() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("M,N_out,K_in", [(131072, 512, 512), (128 * 148 + 77, 256, 512), (40000, 512, 256)])
def test_dgrad_cta_pair_path(M, N_out, K_in):
    """M >= 18 944 and k_in % 256 == 0: the CTA-pair kernel (tcgen05 cta_group::2) with the tanh' operand fetched
    by TMA and the bias gradient of the layer below (column sums of the bf16 output) from the epilogue."""
    from rsoccer_isaac_cleanrl_b200.engine import EPI_DTANH_BF16, gemm_bf16
    dz, wt = _rand((M, N_out), 15), _rand((K_in, N_out), 16, N_out ** -0.5)
    y = torch.tanh(_rand((M, K_in), 17).float()).to(torch.bfloat16)
    out = torch.empty((M, K_in), device="cuda", dtype=torch.bfloat16)
    gemm_bf16(dz, wt, out, EPI_DTANH_BF16, aux=y)
    ref = (dz.float() @ wt.float().t()) * (1 - y.float() ** 2)
    assert (out.float() - ref).abs().max().item() < 3e-2
    out2 = torch.empty_like(out)
    colsum = torch.zeros(K_in, device="cuda")
    gemm_bf16(dz, wt, out2, EPI_DTANH_BF16, aux=y, colsum=colsum)
    assert torch.equal(out, out2)
    want = out.double().sum(0)
    assert (colsum.double() - want).abs().max().item() < 2e-3 * want.abs().max().item() + 1e-2


# (split-K wgrad, MN-major operands, 128 x 256 / 128 x 64 tiles, vector reductions into dW: ragged batch, one and two
# column tiles, the minibatch-sized reductions)
@pytest.mark.parametrize("batch,N_out,K_in,splits", [(4096, 256, 64, 8), (8192, 512, 256, 16), (4000, 512, 512, 7),
                                                     (131072, 256, 512, 32), (131072, 512, 512, 18), (40003, 512, 256, 9),
                                                     (131072, 256, 64, 49)])
def test_wgrad_mn_major_split_k(batch, N_out, K_in, splits):
    from rsoccer_isaac_cleanrl_b200.engine import EPI_ATOMIC_F32, gemm_bf16
    dz, x = _rand((batch, N_out), 8, 0.1), _rand((batch, K_in), 9)
    dw = torch.zeros((N_out, K_in), device="cuda")
    gemm_bf16(dz, x, dw, EPI_ATOMIC_F32, splits=splits, mn_major=True)
    ref = dz.float().t() @ x.float()
    scale = ref.abs().max().item()
    assert (dw - ref).abs().max().item() < 2e-3 * scale + 1e-3, ((dw - ref).abs().max().item(), scale)


def test_split_k_k_major_atomic():
    from rsoccer_isaac_cleanrl_b200.engine import EPI_ATOMIC_F32, gemm_bf16
    a, b = _rand((256, 1024), 10), _rand((128, 1024), 11, 1 / 32)
    out = torch.zeros((256, 128), device="cuda")
    gemm_bf16(a, b, out, EPI_ATOMIC_F32, splits=4)
    ref = a.float() @ b.float().t()
    assert (out - ref).abs().max().item() < 5e-3


def test_tc_mlp_matches_fp32_agent_forward_and_backward():
    """The tcgen05 MLP against the reference-shaped fp32 torch Agent (bf16 tolerance)."""
    import types
    from rsoccer_isaac_cleanrl_b200 import ppo
    envs = types.SimpleNamespace(single_observation_space=types.SimpleNamespace(shape=(52,)),
                                 single_action_space=types.SimpleNamespace(shape=(2,)))
    torch.manual_seed(0)
    ref = ppo.Agent(envs, "torch").cuda()
    tc = ppo.Agent(envs, "tc").cuda()
    tc.load_state_dict(ref.state_dict())
    with torch.no_grad():  # make the heads non-trivial (actor head is initialised with std 0.01)
        for a in (ref, tc):
            a.actor_mean[8].weight.mul_(30.0)
    tc.load_state_dict(ref.state_dict())
    M = 4096 + 77
    x = torch.randn(M, 52, device="cuda")
    act = torch.randn(M, 2, device="cuda")
    adv = torch.randn(M, device="cuda")
    outs = []
    for a in (ref, tc):
        a.zero_grad()
        _, logp, ent, v = a.get_action_and_value(x, act)
        loss = (-(adv * logp).mean()) + 0.5 * (v.view(-1) ** 2).mean() - 0.01 * ent.mean()
        loss.backward()
        outs.append((logp.detach(), v.detach(), {k: p.grad.clone() for k, p in a.named_parameters()}))
    (lp0, v0, g0), (lp1, v1, g1) = outs
    assert (v0 - v1).abs().max().item() < 0.05 * v0.abs().max().item() + 0.02
    assert (lp0 - lp1).abs().max().item() < 0.05 * lp0.abs().max().item() + 0.05
    for k in g0:
        num = (g0[k] - g1[k]).norm().item()
        den = g0[k].norm().item() + 1e-12
        assert num / den < 0.06, (k, num / den)


def _mlp_pair(n_act, seed):
    """Actor-shaped and critic-shaped tanh MLPs with non-trivial heads, plus their bf16 operand copies."""
    from rsoccer_isaac_cleanrl_b200.tc_mlp import MlpWeights
    torch.manual_seed(seed)
    def seq(n_out):
        dims = [52, 256, 512, 512, 256]
        layers = []
        for i in range(4):
            layers += [torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.Tanh()]
        layers.append(torch.nn.Linear(256, n_out))
        s = torch.nn.Sequential(*layers).cuda()
        with torch.no_grad():
            for m in s:
                if isinstance(m, torch.nn.Linear):
                    m.weight.mul_(2.0); m.bias.normal_(0, 0.3)
        return s
    return MlpWeights(seq(n_act)), MlpWeights(seq(1))


@pytest.mark.parametrize("M,n_act,ew", [(4096, 2, 8), (4096, 2, 4), (1000, 6, 8), (128, 2, 8), (77, 6, 4), (65535, 2, 8),
                                        (16384, 6, 8)])
def test_fused_mlp_forward_matches_the_layer_by_layer_path(M, n_act, ew):
    """include/vss_b200.h vss_mlp_forward_fused (csrc/mlp_fused.cu) against four vss_gemm_bf16_tn launches + the head
    kernel: the hidden activations are the same bits (same accumulation order, same epilogue arithmetic), the 256-term
    head sum only differs in summation order -> |diff| <= 2e-5 * (sum |w| + |out|); and both against fp32 torch on
    the bf16-rounded weights (bf16 activations: 3e-2 of the output scale)."""
    from rsoccer_isaac_cleanrl_b200.engine import gather_pad_bf16, mlp_forward_fused
    from rsoccer_isaac_cleanrl_b200.tc_mlp import forward_explicit
    actor, critic = _mlp_pair(n_act, seed=M % 13)
    x = torch.randn(M, 52, device="cuda")
    x16 = gather_pad_bf16(x, None, 64)
    ref_a, _ = forward_explicit(actor, x16)
    ref_c, _ = forward_explicit(critic, x16)
    nets = [(mw.w16, [b.detach() for b in mw.bs], mw.head_w.detach(), mw.head_b.detach(), None) for mw in (actor, critic)]
    out_a, out_c = mlp_forward_fused(x16, nets, epilogue_warps=ew)
    torch.cuda.synchronize()
    for out, ref, mw in ((out_a, ref_a, actor), (out_c, ref_c, critic)):
        assert out.shape == ref.shape and torch.isfinite(out).all()
        bound = 2e-5 * (mw.head_w.detach().abs().sum(1).max().item() + ref.abs().max().item())
        assert (out - ref).abs().max().item() <= bound, ((out - ref).abs().max().item(), bound)
        with torch.no_grad():
            h = x16[:, :52].float()
            for l in range(4):
                h = torch.tanh(h @ mw.w16[l][:, :h.shape[1]].float().t() + mw.bs[l])
            fp32 = h @ mw.head_w.t() + mw.head_b
        assert (out - fp32).abs().max().item() < 3e-2 * fp32.abs().max().item() + 1e-2
    # one network alone, and a second call (barrier phases, TMEM release) give the same numbers
    only_c, = mlp_forward_fused(x16, nets[1:], epilogue_warps=ew)
    again_a, again_c = mlp_forward_fused(x16, nets, epilogue_warps=ew)
    assert torch.equal(only_c, out_c) and torch.equal(again_a, out_a) and torch.equal(again_c, out_c)


@pytest.mark.parametrize("M,n_act", [(4096, 2), (1000, 6)])
def test_fused_mlp_forward_samples_like_policy_sample(M, n_act):
    """vss_mlp_sampling: the action and log-prob drawn inside the fused launch are those of vss_policy_sample on the
    same mean with call index counter + call_offset (same arithmetic, same Philox stream); the counter is not advanced;
    the mean need not be stored."""
    from rsoccer_isaac_cleanrl_b200.engine import gather_pad_bf16, mlp_forward_fused, policy_sample
    actor, critic = _mlp_pair(n_act, seed=3)
    x16 = gather_pad_bf16(torch.randn(M, 52, device="cuda"), None, 64)
    nets = [(mw.w16, [b.detach() for b in mw.bs], mw.head_w.detach(), mw.head_b.detach()) for mw in (actor, critic)]
    logstd = torch.linspace(-0.7, 0.2, n_act, device="cuda")
    ctr = torch.full((1,), 5, device="cuda", dtype=torch.int32)
    act, lp = torch.empty((M, n_act), device="cuda"), torch.empty(M, device="cuda")
    samp = dict(logstd=logstd, counter=ctr, seed=1234567, call_offset=3, action=act, logprob=lp)
    mean, value = mlp_forward_fused(x16, [nets[0] + (None,), nets[1] + (None,)], sampling=samp)
    assert int(ctr.item()) == 5
    ctr2 = torch.full((1,), 8, device="cuda", dtype=torch.int32)
    act_ref, lp_ref = policy_sample(mean, logstd, 1234567, ctr2)
    assert int(ctr2.item()) == 9
    assert torch.equal(act, act_ref) and torch.equal(lp, lp_ref)
    # without storing the mean: same draws, same value
    act2, lp2 = torch.empty_like(act), torch.empty_like(lp)
    samp.update(action=act2, logprob=lp2)
    none, value2 = mlp_forward_fused(x16, [nets[0] + (False,), nets[1] + (None,)], sampling=samp)
    assert none is None and torch.equal(act2, act) and torch.equal(lp2, lp) and torch.equal(value2, value)
    # the actor alone (how the rollout launches it when the critic runs on the side stream): same draws again
    act3, lp3 = torch.empty_like(act), torch.empty_like(lp)
    samp.update(action=act3, logprob=lp3)
    mlp_forward_fused(x16, [nets[0] + (False,)], sampling=samp)
    assert torch.equal(act3, act) and torch.equal(lp3, lp)
    samp.update(action=act2, logprob=lp2)
    # a different call index gives different noise; the statistics are those of N(mean, exp(logstd))
    samp.update(call_offset=4)
    mlp_forward_fused(x16, [nets[0] + (False,), nets[1] + (None,)], sampling=samp)
    assert not torch.equal(act2, act)
    zs = (act - mean) / logstd.exp()
    assert abs(zs.mean().item()) < 5 / (M * n_act) ** 0.5 and abs(zs.std().item() - 1) < 5 / (2 * M * n_act) ** 0.5


def test_fused_mlp_forward_is_bit_stable_under_load():
    """The barrier protocol of csrc/mlp_fused.cu (TMEM halves, per-block ready / free mbarriers, generic -> async proxy
    fences) under timing perturbation: 60 launches over several waves of CTAs, while another stream keeps the SMs and
    the L2 busy with large GEMMs and copies, give the same bits every time, for both epilogue shapes."""
    from rsoccer_isaac_cleanrl_b200.engine import gather_pad_bf16, mlp_forward_fused
    actor, critic = _mlp_pair(6, seed=5)
    M = 40000 + 77
    x16 = gather_pad_bf16(torch.randn(M, 52, device="cuda"), None, 64)
    nets = [(mw.w16, [b.detach() for b in mw.bs], mw.head_w.detach(), mw.head_b.detach(), None) for mw in (actor, critic)]
    refs = {ew: mlp_forward_fused(x16, nets, epilogue_warps=ew) for ew in (8, 4)}
    # (the two shapes sum the head's 256 products in different orders: equal to rounding, not to the bit)
    assert (refs[8][0] - refs[4][0]).abs().max().item() < 1e-4 * refs[8][0].abs().max().item()
    side = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    big = torch.empty(64 << 20, device="cuda")
    for it in range(60):
        with torch.cuda.stream(side):
            if it % 3 == 0:
                a @ a
            elif it % 3 == 1:
                big.copy_(big.flip(0))
        ew = 8 if it % 2 == 0 else 4
        out_a, out_c = mlp_forward_fused(x16, nets, epilogue_warps=ew)
        assert torch.equal(out_a, refs[ew][0]) and torch.equal(out_c, refs[ew][1]), it
    torch.cuda.synchronize()
