"""Pure-Python restatement of the reference's Rust-side post-processing
(/root/reference/src/asr/whisper.rs:9-14, 41-43, 84-128, 175-201) used as the checker for result-level
parity.  Input: raw whisper segments (bytes text, t0, t1, speaker_turn_next)."""
PROMOTIONAL_TEXT = [
    "请不吝点赞", "請不吝點贊", "點贊", "訂閱", "订阅", "打赏", "打賞", "打賞支持明鏡與點點欄目", "打赏支持明镜与点点栏目",
    "並且按下小鈴鐺才能收到最新消息哦!", "請按讚、訂閱、分享!", "明镜需要您的支持 欢迎收看订阅明镜",
    "請按讚,訂閱,分享,打開小鈴鐺,並且按下小鈴鐺才能收到最新消息謝謝觀看",
    "請按讚,訂閱,分享,打開小鈴鐺,並且按下小鈴鐺才能收到最新消息哦!",
]


def add_punctuation(text: str) -> str:                      # whisper.rs:175-201
    if text.endswith(("。", "！", "？", "，")):
        return text
    q = any(k in text for k in ("吗", "呢", "什么", "为何", "怎么"))
    e = any(k in text for k in ("啊", "哇", "太", "真", "好", "真是"))
    return text + ("？" if q else "！" if e else " ")


def is_promotional_text(text: str) -> bool:                 # whisper.rs:41-43
    return any(p in text for p in PROMOTIONAL_TEXT)


def post_process(raw_segments, stream_mode: bool):          # whisper.rs:84-128
    segments, full_text, speaker = [], "", 0
    n = len(raw_segments)
    for i, s in enumerate(raw_segments):
        text = s["text"].decode("utf-8")                    # invalid UTF-8 -> error, like `?` at :85
        if is_promotional_text(text):
            continue
        if i > 0 and raw_segments[i - 1]["speaker_turn_next"]:
            speaker += 1
        processed = add_punctuation(text)
        if stream_mode:
            if i == n - 1:
                segments.append((processed, speaker, float(s["t0"]), float(s["t1"])))
                full_text = processed
        else:
            segments.append((processed, speaker, float(s["t0"]), float(s["t1"])))
            full_text += processed
    return {"segments": segments, "full_text": full_text}
