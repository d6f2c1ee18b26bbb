"""Small end-to-end pass over every kernel family (GPU box); written for `compute-sanitizer --tool memcheck`, which is closed on this pool:
   compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/small_all_paths.py
1 s clips, 2-layer CLIP, one training step at T = 24: everything finishes in seconds outside the sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from avi_talking_b200 import synth, train  # noqa: E402
from avi_talking_b200.smoke import build_models, build_prior  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.cuda.set_device(0)
if which in ("all", "path"):
    for prec in ("bf16", "fp32"):
        m = build_models(prec)
        a = synth.audio(2, 16000, seed=1).cuda()
        emo = torch.randn(2, 24, 30).cuda()
        v = m.predict_from_embeddings(a, emo)
        fv = m.convert_coeff2verts(torch.randn(48, 53).cuda(), 0.1 * torch.randn(48, 6).cuda(), torch.randn(48, 100).cuda())
        out = m.flame(torch.randn(5, 100).cuda(), torch.randn(5, 50).cuda(), 0.1 * torch.randn(5, 6).cuda())
        torch.cuda.synchronize()
        print("path", prec, float(v.abs().max()), float(fv.abs().max()), float(out[0].abs().max()))
if which in ("all", "train"):
    from transformers import Wav2Vec2Config
    from avi_talking_b200.faceformer import FaceformerVert, make_args
    from avi_talking_b200.wav2vec import Wav2Vec2Model
    for prec in ("bf16", "fp32"):
        w2v = Wav2Vec2Model(Wav2Vec2Config(num_hidden_layers=2))
        w2v.load_state_dict(synth.wav2vec2_state(0, layers=2), strict=False)
        template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
        m = FaceformerVert(make_args(feature_dim=64), audio_encoder=w2v, template=template)
        m.load_state_dict(synth.faceformer_state(fd=64, seed=264, variant="vert"), strict=False)
        m.precision = w2v.precision = prec
        m = m.cuda()
        opt = train.FlatAdam(m, lr=1e-4)
        gt = (template + 1e-3 * torch.randn(2, 24, 15069)).cuda()
        audio = synth.audio(2, 16000, seed=3).cuda()
        for _ in range(2):
            opt.zero_grad()
            loss = m.training_loss(audio, gt)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        print("train", prec, float(loss.detach()))
if which in ("all", "prior"):
    from transformers import CLIPTextConfig
    from avi_talking_b200.clip_text import CLIPTextModel
    from avi_talking_b200.diffusion_prior import voxel2style_emb
    cfg = CLIPTextConfig(vocab_size=synth.CLIP_TEXT.vocab, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768)
    for prec in ("bf16", "fp32"):
        c = CLIPTextModel(cfg)
        c.load_state_dict(synth.clip_text_state(60, 2), strict=False)
        c.precision = prec
        c = c.cuda().eval()
        prior = build_prior(prec)
        inp = synth.prior_inputs(4, 8)
        s = voxel2style_emb(c.text_to_voxel(synth.clip_tokens(4, seed=1).cuda()), prior, timesteps_prior=8, image_embed=inp["image_embed"].cuda(),
                            noise=inp["noises"][:7].cuda())
        torch.cuda.synchronize()
        print("prior", prec, float(s.abs().max()))
print("sanitize_small done")
