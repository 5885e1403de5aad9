"""INI -> typed `Conf`, the config surface of the DMT hot path.

Mirrors `DMT_code/conf/recsys_conf.py` (class `Conf`, :17-366) for every option
the forward/backward ranking path reads, including its typing quirks:

* `reset(section, option, fn, default)` converts in place and silently falls
  back to `default` on *any* failure (recsys_conf.py:234-242);
* `zero_pad` is never converted, so any non-empty string -- even "false" -- is
  truthy (recsys_conf.py:158);
* `learning_rate` / `step_boundary` are comma lists (recsys_conf.py:93-94);
* class weights become lists ordered by ascending label (util.py:132-144).

Deliberate differences (all outside the hot path): no `os.makedirs` of the model
directory unless `create_dirs=True` (recsys_conf.py:112-113 does it
unconditionally), no `hadoop` shell-outs (recsys_conf.py:340-347), and the label
statistics file is only read when it exists.
"""
import configparser
import os

from . import keys as K


class Conf:
    def __init__(self, conf_path="./", conf_file="dnn.model.conf", create_dirs=False,
                 overrides=None):
        parser = configparser.ConfigParser()
        read_ok = parser.read(os.path.join(conf_path, conf_file) if not conf_path.endswith("/")
                              else conf_path + conf_file)
        if not read_ok:
            raise FileNotFoundError("config file not found: %s%s" % (conf_path, conf_file))
        self.conf_parser = parser
        self.tag = self.get_tag(conf_file)
        self.label_cnt_lst = []
        self.yes_label_cnt_lst = []
        self.conf_sections = {s: dict(parser.items(s)) for s in parser.sections()}
        # test / bench hook: {(section, option): raw string} applied before typing,
        # equivalent to editing the INI file.
        for (sec, opt), raw in (overrides or {}).items():
            self.conf_sections.setdefault(sec, {})[opt] = raw

        R = self.reset
        R(K.PARAMETER, K.LABEL_WEIGHT, K.csv_to_int_list, None)
        R(K.PARAMETER, K.LOSS_WEIGHT, K.csv_to_float_list, None)
        R(K.EXPORT_MODEL, K.EXPORT_WEIGHT, K.csv_to_float_list, None)
        for opt in (K.FEAT_DIM, K.OUTPUT_UNITS, K.num_experts, K.EPOCH_NUM, K.BATCH_SIZE,
                    K.TEST_BATCH_SIZE, K.VALIDATION_BATCH_SIZE, K.VALIDATE_STEP,
                    K.MAX_ITER_STEP, K.TOTAL_EXAMPLE_NUM):
            R(K.MODEL, opt, int, None)
        for opt in (K.HIDDEN_UNITS, K.HIDDEN_UNITS_BIAS, K.hidden_units_bottom,
                    K.hidden_units_task, K.FILTER_SHAPE):
            R(K.MODEL, opt, K.csv_to_int_list, None)
        for opt in (K.DROPOUT, K.DROPOUT_BOTTOM, K.DROPOUT_TASK):
            R(K.MODEL, opt, K.csv_to_float_list, None)
        R(K.MODEL, K.IS_USE_FEATURE, K.str_to_bool, True)
        R(K.MODEL, K.SHUFFLE_SIZE, int, 100000)
        R(K.MODEL, K.IS_BN, K.str_to_bool, None)
        R(K.MODEL, K.BN_DECAY, float, 0.999)
        R(K.MODEL, K.IS_DROPOUT, K.str_to_bool, None)
        R(K.MODEL, K.LOSS_CTR_REL_METHOD, str, None)
        R(K.MODEL, K.propensity_em, K.str_to_bool, False)
        R(K.MODEL, K.propensity_em_type, str, None)
        R(K.EMBEDDING, K.attention_embed_seq_ts, str, "")

        self.labels = self.get_labels(self[K.CLASS_WEIGHT][K.TRAIN_WEIGHT])
        for opt in (K.TRAIN_WEIGHT, K.VALID_WEIGHT, K.WEIGHT_CTR, K.WEIGHT_ECVR):
            R(K.CLASS_WEIGHT, opt, K.parse_weight, None)
        R(K.MODEL, K.WND_WD, float, None)
        R(K.MODEL, K.L2_EMB_LAMBDA, float, None)
        R(K.MODEL, K.ENABLE_SSP, K.str_to_bool, True)

        model = self[K.MODEL]
        model[K.STEP_BOUNDARY] = [int(s) for s in model[K.STEP_BOUNDARY].split(",")]
        model[K.LEARNING_RATE] = [float(s) for s in model[K.LEARNING_RATE].split(",")]

        path = self.conf_sections.setdefault(K.PATH, {})
        out = path.get(K.OUTPUT_PATH, "./")
        path[K.MODEL_PATH] = out + self.tag + ".model/"
        path[K.MODEL_FROZEN_PATH] = path[K.MODEL_PATH] + "frozen/"
        path[K.MODEL_IMP_PATH] = path[K.MODEL_PATH] + "imp/"
        path[K.VALIDATION_RESULT] = out + self.tag + ".validation.result"
        path[K.TRAIN_RESULT] = out + self.tag + ".train.result"
        if create_dirs:
            os.makedirs(os.path.expanduser(path[K.MODEL_PATH]), exist_ok=True)
        for opt in (K.TRAIN_DATA_PATH, K.TEST_DATA_PATH, K.VALIDATION_DATA_PATH):
            if path.get(opt) and path[opt][-1] != "/":
                path[opt] += "/"

        emb = self[K.EMBEDDING]
        self.embedding_list = self.get_emb(emb.get(K.EMB, ""))
        self.embedding_list_bias = self.get_emb(emb.get(K.EMB_BIAS, ""))
        self.attention_embed_pairs = self.get_attention_embed_v2(emb.get(K.ATTENTION_EMBED, ""))
        self.attention_embed_seq_ts = self.get_attention_embed_ts(emb[K.attention_embed_seq_ts])
        self.sim_embed = self.get_attention_embed(emb.get(K.SIM_EMBED, ""))
        self.embedding_init_info = self.get_emb_init_info(emb.get(K.UPDATE_EMB, ""))

        self.weight_ctr = self[K.CLASS_WEIGHT][K.WEIGHT_CTR]
        self.weight_ecvr = self[K.CLASS_WEIGHT][K.WEIGHT_ECVR]

        if K.SCHEMA in self.conf_sections and K.HEADER_SCHEMA in self[K.SCHEMA]:
            self[K.SCHEMA][K.HEADER_SCHEMA] = [
                s.strip() for s in self[K.SCHEMA][K.HEADER_SCHEMA].split(",")]

        stat = os.path.expanduser(path.get(K.TRAIN_DATA_STAT_PATH) or "")
        if stat and os.path.isfile(stat):
            self._apply_label_stats(stat)

        self.model_type = model[K.MODEL_TYPE]
        self.zero_pad = model.get(K.zero_pad, "")          # raw string: truthy if non-empty
        self.is_unbias_model = "unbias" in self.model_type
        self.propensity_em = model[K.propensity_em]
        self.propensity_em_type = model[K.propensity_em_type]
        if self.is_unbias_model:
            self.loss_unbias_method = model.get(K.loss_unbias_method)
            R(K.MODEL, K.dropout_rate_bias, K.csv_to_float_list, None)
            self.dropout_rate_bias = model[K.dropout_rate_bias]
            self.loss_ctr_rel_method = model[K.LOSS_CTR_REL_METHOD]
        self.is_use_feature = model[K.IS_USE_FEATURE]

        if "transformer" in self.model_type:
            for opt in (K.transformer_d_model, K.transformer_d_ff, K.transformer_num_heads,
                        K.transformer_num_blocks_encode, K.transformer_num_blocks_decode,
                        K.transformer_maxlen_k, K.transformer_maxlen_q):
                R(K.MODEL, opt, int, None)
            R(K.MODEL, K.transformer_dropout_rate, float, None)
            R(K.MODEL, K.transformer_is_trans_input_by_mlp, K.str_to_bool, False)
            R(K.MODEL, K.transformer_position_encoding_method, str, "position_sin_cos")
            R(K.MODEL, K.transformer_is_trans_out_concat_item, K.str_to_bool, True)
            R(K.MODEL, K.transformer_is_trans_out_by_mlp, K.str_to_bool, False)
            R(K.MODEL, K.transformer_is_decoder_add_pos_emb, K.str_to_bool, False)
            self.d_model = model[K.transformer_d_model]
            self.d_ff = model[K.transformer_d_ff]
            self.num_heads = model[K.transformer_num_heads]
            self.num_blocks_encode = model[K.transformer_num_blocks_encode]
            self.num_blocks_decode = model[K.transformer_num_blocks_decode]
            self.maxlen_k = model[K.transformer_maxlen_k]
            self.maxlen_q = model[K.transformer_maxlen_q]
            self.dropout_rate = model[K.transformer_dropout_rate]
            self.is_trans_input_by_mlp = model[K.transformer_is_trans_input_by_mlp]
            self.position_encoding_method = model[K.transformer_position_encoding_method]
            self.is_use_seq_ts = len(emb[K.attention_embed_seq_ts]) >= 1
            self.is_trans_out_concat_item = model[K.transformer_is_trans_out_concat_item]
            self.is_trans_out_by_mlp = model[K.transformer_is_trans_out_by_mlp]
            self.is_decoder_add_pos_emb = model[K.transformer_is_decoder_add_pos_emb]

    # -- typing helper (recsys_conf.py:234-242) --------------------------------
    def reset(self, section, option, fn, default):
        try:
            self.conf_sections[section][option] = fn(self.conf_sections[section][option])
        except Exception:
            self.conf_sections.setdefault(section, {})[option] = default

    def __getitem__(self, section):
        return self.conf_sections[section]

    def get_conf(self, section, option=None):
        try:
            sec = self.conf_sections[section]
            return sec if option is None else sec[option]
        except KeyError:
            return None

    @staticmethod
    def get_tag(conf_file):
        """`a.b.conf` -> `a.b` (recsys_conf.py:260-265)."""
        return conf_file[:-len(".conf")] if conf_file.split(".")[-1] == "conf" else conf_file

    # -- [embedding] grammar (recsys_conf.py:269-338) ---------------------------
    @staticmethod
    def get_emb(spec):
        """`Name:V:D:feature:side#...` -> [[Name, V, D, feature, side], ...]."""
        if len(spec) <= 2:
            return []
        out = []
        for entry in spec.split("#"):
            f = entry.split(":")
            f[1], f[2] = int(f[1]), int(f[2])
            out.append(f)
        return out

    @staticmethod
    def get_attention_embed(spec):
        if len(spec) <= 2:
            return []
        return [tuple(p.split(":")[:2]) for p in spec.split("#")]

    @staticmethod
    def get_attention_embed_v2(spec):
        """`u:i#u:i|u:i#...` -> one pair list per behaviour sequence."""
        if len(spec) <= 2:
            return []
        return [[tuple(p.split(":")[:2]) for p in seq.split("#")] for seq in spec.split("|")]

    @staticmethod
    def get_attention_embed_ts(spec):
        if len(spec) <= 1:
            return []
        return [s.strip() for s in spec.split("|")]

    @staticmethod
    def get_emb_init_info(spec):
        out = {}
        for entry in spec.split("#"):
            f = entry.split(":")
            if len(f) == 2:
                out[f[0]] = f[1]
        return out

    def get_idschema(self):
        return [e[3] for e in self.embedding_list]

    def get_idschema_bias(self):
        return [e[3] for e in self.embedding_list_bias]

    @staticmethod
    def get_labels(label_weight_str):
        return sorted(int(item.strip().split(":")[0]) for item in label_weight_str.split(","))

    def _apply_label_stats(self, stat_file):
        """Label counts cap `max_iter_step` at epochs*N/(B*n_gpu) (recsys_conf.py:139-151)."""
        model = self[K.MODEL]
        with open(stat_file) as fh:
            self.label_cnt_lst = [int(line.strip()) for line in fh if line.strip()]
        model[K.TOTAL_EXAMPLE_NUM] = sum(self.label_cnt_lst)
        self.label_cnt_lst = [x / self.label_cnt_lst[-1] for x in self.label_cnt_lst]
        n_gpu = len(model[K.GPU_VISIBLE].split(","))
        total_step = int(model[K.EPOCH_NUM] * model[K.TOTAL_EXAMPLE_NUM] / (model[K.BATCH_SIZE] * n_gpu))
        model[K.SAVE_CKPT_NUMS] = int(model[K.MAX_ITER_STEP] / model[K.VALIDATE_STEP])
        if model[K.MAX_ITER_STEP] > total_step:
            model[K.MAX_ITER_STEP] = total_step
            model[K.SAVE_CKPT_NUMS] = int(total_step * n_gpu / model[K.VALIDATE_STEP])
