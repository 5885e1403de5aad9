"""Command-line flags of the reference driver (`DMT_code/parse/parse.py:4-49`): same names,
defaults and string typing (every flag is a string, `'false'`/`'true'` included)."""
import argparse

_FLAGS = (
    ("conf_path", "./conf/settings/", "config directory"),
    ("conf_file", "demo.conf", "config file; its name minus .conf is the run tag"),
    ("model_ckpt", "model.ckpt-0", "checkpoint name; the global step is parsed from its suffix"),
    ("is_train", "false", "'true' selects training"),
    ("is_valid", "false", "'true' selects validation"),
    ("test_tag", "clk", "clk or ord"),
    ("test_score_method", "ctr", "rel (sigmoid(logit)) or ctr (sigmoid(logit + bias))"),
    ("is_test", "false", "'true' selects prediction"),
)


def argument_parse(argv=None):
    parser = argparse.ArgumentParser(description="dnn conf")
    for name, default, text in _FLAGS:
        parser.add_argument("--" + name, dest=name, default=default, help=text)
    return vars(parser.parse_args(argv))
