"""Caption templates for use_templates=True: same nine strings, same order, as reference
clip_text_templates.py (they are data the score definition depends on, clip_scoring.py:156-162)."""
_PREFIXES = ['', 'a photo of ', 'a bad photo of ', 'a good photo of ', 'a low resolution photo of ',
             'a cropped photo of ', 'a bright photo of ', 'a dark photo of ', 'a painting of ']
CLIP_TEMPLATES = [p + '{}' for p in _PREFIXES]
