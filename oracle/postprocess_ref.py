"""CPU restatement of the reference post-processing -- TEST INFRASTRUCTURE ONLY (the product is
advancedliteratemachinery_b200/csrc/postproc.cu; nothing under advancedliteratemachinery_b200/ imports this).

  omni_results(...)   OCR/OmniParser/engine/val.py:70-100 + utils/misc.py:147-189
  mgp_fuse_ref(...)   OCR/MGP-STR/test_final.py:176-240 (the inline fusion block of `validation`)

Pinned by oracle/gen_golden_post.py: the OmniParser half against the reference's own `decode_pred_seq` (imported,
unmodified), the MGP-STR strings against the reference's `TokenLabelConverter.{char,bpe,wp}_decode` driven by real
HuggingFace tokenizers over synthetic vocabularies; the fusion block is inline code in the reference (not callable),
so it is restated here line by line.
"""
import torch


def omni_results(index_seqs, prob_seqs, file_name, orig_size, args):
    """val.py:70-100.  index_seqs = [pt [1,2N], poly [1,32N], rec [1,N,L]] int64, prob_seqs [N,L] f32."""
    pt = index_seqs[0].reshape(-1)
    assert len(pt) % 2 == 0, 'odd-length point sequences crash the reference too (val.py:73-74 keeps seq[:1])'
    pt = pt.reshape(-1, 2)
    poly = index_seqs[1].reshape(-1, 32)
    rec = index_seqs[2].reshape(-1, args.rec_length)
    probs = prob_seqs.reshape(-1, args.rec_length)
    image_h, image_w = torch.as_tensor(orig_size)          # target['orig_size'] is an int64 tensor (val.py:85)
    out = []
    for i in range(pt.shape[0]):
        px = (pt[i, 0] / args.num_bins).item() * image_w   # misc.py:152-154, val.py:88-89 (float32 tensor math)
        py = (pt[i, 1] / args.num_bins).item() * image_h
        polygon = (poly[i] / args.num_bins) * torch.tensor([image_w, image_h] * 16)   # misc.py:159, val.py:90
        chars, kept = [], []
        for j in range(args.rec_length):                   # misc.py:169-179
            t = int(rec[i, j])
            if t == args.recog_pad_index or t == args.rec_eos_index:
                break
            if t == args.recog_pad_index - 1:
                continue
            chars.append(args.chars[t - args.num_bins])
            kept.append(probs[i, j].item())
        out.append({'image_id': file_name, 'pts': [[px.item(), py.item()]],
                    'score': sum(kept) / (len(kept) + 1e-5), 'polys': polygon.reshape(-1, 2).tolist(),
                    'rec': ''.join(chars)})
    return out


def kie_walk(tokens, probs, inst_pos, poly, rec, orig_size, args, classes, class_base):
    """decode_vie_pt_poly_rec_seq (transformer.py:148-215) with the per-instance poly / rec decode loops replaced by the
    already decoded ids (what the batched GPU loops return): tokens / probs of the point sequence, inst_pos = start index
    of every (x, y) pair, poly [n,32], rec [n,L]."""
    image_h, image_w = int(orig_size[0]), int(orig_size[1])
    result, words, rects = [], [], []
    by_pos = {int(p): k for k, p in enumerate(inst_pos)}
    i = 0
    while i < len(tokens):
        if tokens[i] < args.num_bins:
            if i + 1 <= len(tokens) - 1 and tokens[i + 1] < args.num_bins:
                k = by_pos[i]
                pts = torch.as_tensor(poly[k]).reshape(-1, 2)
                rects.append([image_w * pts[:, 0].min().item() / args.num_bins, image_h * pts[:, 1].min().item() / args.num_bins,
                              image_w * pts[:, 0].max().item() / args.num_bins, image_h * pts[:, 1].max().item() / args.num_bins])
                chars = []
                for t in torch.as_tensor(rec[k]).tolist():
                    if t == args.recog_pad_index or t == args.rec_eos_index:
                        break
                    if t == args.recog_pad_index - 1:
                        continue
                    chars.append(args.chars[t - args.num_bins])
                words.append(''.join(chars))
                i += 2
            else:
                i += 1
        else:
            result.append((' '.join(words), classes[tokens[i] - class_base], torch.as_tensor(probs)[i].item(), rects))
            words, rects = [], []
            i += 1
    return result


def mgp_fuse_ref(strings, ids, prob):
    """test_final.py:176-240 for one batch.  strings: [3][B] the converter's decoded strings (char / bpe / wp);
    ids, prob: [3, B, T] incl. position 0.  Returns texts [3][B], conf [3][B] (float), fused [B], source [B]."""
    B = ids.shape[1]
    eos_str = ['[s]', '#', '[SEP]']
    eos_id = [None, 2, 102]
    texts = [[None] * B for _ in range(3)]
    conf = [[0.0] * B for _ in range(3)]
    fused, source = [], []
    for b in range(B):
        best, out_pred, src = 0.0, None, -1
        for hd in range(3):
            pred = strings[hd][b]
            max_prob = prob[hd, b, 1:]
            eos = pred.find(eos_str[hd])
            pred = pred[:eos]
            if hd == 0:
                sl = max_prob[:eos + 1]                    # :188 (string index)
            else:
                lst = ids[hd, b, 1:].tolist()
                try:
                    k = lst.index(eos_id[hd])              # :202-205 / :221-224
                except ValueError:
                    k = -1
                sl = max_prob[:k + 1]
            try:
                c = sl.cumprod(dim=0)[-1]
            except Exception:
                c = 0.0
            texts[hd][b] = pred
            conf[hd][b] = float(c)
            if c > best:
                best, out_pred, src = c, pred, hd
        fused.append(out_pred if out_pred is not None else '')
        source.append(src)
    return texts, conf, fused, source
