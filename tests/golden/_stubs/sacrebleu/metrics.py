class _BLEU:
    TOKENIZERS = ['none', 'zh', '13a', 'intl', 'char', 'ja-mecab', 'ko-mecab', 'spm', 'flores101', 'flores200']


METRICS = {'BLEU': _BLEU, 'CHRF': object, 'TER': object}
BLEU = _BLEU
