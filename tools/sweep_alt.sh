python tools/sweep_variant.py alt
python tools/sweep_variant.py noalt '{"*": {"alternate": 0}}'
python tools/sweep_variant.py alt2
python tools/sweep_variant.py noalt2 '{"*": {"alternate": 0}}'
