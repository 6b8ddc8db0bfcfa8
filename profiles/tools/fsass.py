import sys,subprocess,re
lib,name=sys.argv[1],sys.argv[2]
out=subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
parts=re.split(r'\n\s*Function : ',out)
for p in parts[1:]:
    fn=p.split('\n',1)[0].strip()
    if fn==name:
        print(p); break
