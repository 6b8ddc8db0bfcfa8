import pandas as pd, io, sys
f=sys.argv[1]
lines=open(f).read().split('\n')
df=pd.read_csv(io.StringIO('\n'.join(lines[1:])))
df['op']=df['Source'].str.strip().str.replace(r'^@!?U?P\d+\s+','',regex=True).str.split().str[0].str.split('.').str[0]
ie='Instructions Executed'; sm='# Samples'
tot=df[ie].sum(); ts=df[sm].sum()
print('total inst',tot,'samples',ts, 'n sass', len(df))
g=df.groupby('op').agg(inst=(ie,'sum'),samp=(sm,'sum'),thr=('Thread Instructions Executed','sum')).sort_values('inst',ascending=False)
g['inst%']=100*g.inst/tot; g['samp%']=100*g.samp/ts; g['lanes']=g.thr/g.inst
print(g.head(30).to_string())
df['lanes']=df['Avg. Threads Executed']
for lo,hi in [(0,8),(8,16),(16,24),(24,30),(30,33)]:
    m=(df.lanes>=lo)&(df.lanes<hi)
    print(lo,hi,'inst%',100*df[ie][m].sum()/tot,'samp%',100*df[sm][m].sum()/ts)
st=[c for c in df.columns if c.startswith('stall_') and 'Not Issued' not in c]
print((df[st].sum()/df[st].sum().sum()*100).sort_values(ascending=False).head(10))
