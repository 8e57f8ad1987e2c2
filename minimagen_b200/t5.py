"""Text-encoder shim (reference: minimagen/t5.py).  The frozen T5 encoder is OUT OF SCOPE of this package (SURVEY.md
section 2, row 6): it runs once per prompt batch, not per denoising step.  Only the embedding-dimension table that
`Unet`'s default argument needs is kept; `t5_encode_text` defers to HuggingFace transformers when it is importable
and the checkpoint is cached locally."""
import torch

MAX_LENGTH = 256
DEFAULT_T5_NAME = 't5_base'

_T5_DIMS = {'t5_small': 512, 't5_base': 768, 't5_large': 1024, 't5_3b': 1024, 't5_11b': 1024,
            'small': 512, 'base': 768, 'large': 1024, '3b': 1024, '11b': 1024}


def get_encoded_dim(name):
    """reference: t5.py:87-90"""
    return _T5_DIMS[name]


def t5_encode_text(text, name='t5_base', max_length=MAX_LENGTH):
    """reference: t5.py:31-84 -- (embeddings [b, L, D] with padded positions zeroed, bool mask [b, L])."""
    try:
        from transformers import T5EncoderModel, T5Tokenizer
    except Exception as e:   # pragma: no cover
        raise RuntimeError("t5_encode_text needs `transformers`; pass text_embeds/text_masks instead") from e
    hf = name if name.startswith('t5') and '_' not in name else name.replace('_', '-')
    tokenizer = T5Tokenizer.from_pretrained(hf)
    model = T5EncoderModel.from_pretrained(hf).eval()
    device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    model = model.to(device)
    enc = tokenizer(text, padding='longest', max_length=max_length, truncation=True, return_tensors='pt')
    ids, mask = enc.input_ids.to(device), enc.attention_mask.to(device)
    with torch.no_grad():
        emb = model(input_ids=ids, attention_mask=mask).last_hidden_state
    mask = mask.bool()
    emb = emb.masked_fill(~mask[..., None], 0.)
    return emb, mask
