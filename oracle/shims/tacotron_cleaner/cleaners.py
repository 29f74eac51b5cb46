def custom_english_cleaners(x): return x
